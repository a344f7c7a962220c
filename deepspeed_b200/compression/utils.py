"""Straight-through quantisers / binarisers (reference ``compression/utils.py``)."""
import torch
from torch import autograd


class TopKBinarizer(autograd.Function):
    """mask = 1 for the top ``threshold`` fraction of scores; gradient passes straight through to the scores."""

    @staticmethod
    def forward(ctx, inputs, threshold, sigmoid):
        if sigmoid:
            threshold = torch.sigmoid(threshold).item()
        ctx.sigmoid = sigmoid
        mask = torch.zeros_like(inputs)
        k = int(threshold * inputs.numel())
        if k > 0:
            idx = inputs.flatten().topk(k).indices
            mask.view(-1)[idx] = 1.0
        ctx.save_for_backward(mask)
        return mask

    @staticmethod
    def backward(ctx, grad_output):
        (mask, ) = ctx.saved_tensors
        if ctx.sigmoid:
            return grad_output.clone(), ((grad_output * mask).sum()).view(-1), None
        return grad_output.clone(), None, None


class SymQuantizer(autograd.Function):

    @staticmethod
    def forward(ctx, input, num_bits, min_value=None, max_value=None, num_groups=1):
        assert (min_value is None and max_value is None) or (min_value is not None and max_value is not None
                                                             and num_groups == 1)
        q_range = 2**num_bits
        shape = input.shape
        x = input.reshape(num_groups, -1)
        if min_value is None:
            max_in = x.abs().amax(dim=-1, keepdim=True)
        else:
            max_in = torch.max(min_value.abs(), max_value).view(-1)
        scale = 2 * max_in / q_range
        scale = torch.where(scale == 0, torch.ones_like(scale), scale)
        out = (x / scale).round().clamp(-q_range // 2, q_range // 2 - 1) * scale
        return out.reshape(shape).contiguous()

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output.clone(), None, None, None, None


class AsymQuantizer(autograd.Function):

    @staticmethod
    def forward(ctx, input, num_bits, min_value=None, max_value=None, num_groups=1):
        q_range = 2**num_bits
        shape = input.shape
        x = input.reshape(num_groups, -1)
        if min_value is None:
            mn, mx = x.amin(dim=-1, keepdim=True), x.amax(dim=-1, keepdim=True)
        else:
            mn, mx = min_value, max_value
        scale = (mx - mn) / q_range
        scale = torch.where(scale == 0, torch.ones_like(scale), scale)
        zero = (mn / scale).round() * scale
        out = ((x - zero) / scale).round().clamp(0, q_range - 1) * scale + zero
        return out.reshape(shape).contiguous()

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output.clone(), None, None, None, None


class TernaryQuantizer(autograd.Function):

    @staticmethod
    def forward(ctx, input, num_bits, min_value=None, max_value=None, num_groups=1):
        assert min_value is None and max_value is None
        x = input.reshape(num_groups, -1)
        n = x.shape[1]
        m = x.norm(p=1, dim=1).div(n)
        thres = (0.7 * m).view(-1, 1)
        pos, neg = (x > thres).type(input.type()), (x < -thres).type(input.type())
        mask = (x.abs() > thres).type(input.type())
        alpha = ((mask * x).abs().sum(dim=1) / mask.sum(dim=1).clamp(min=1)).view(-1, 1)
        return (alpha * pos - alpha * neg).reshape(input.shape).contiguous()

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output.clone(), None, None, None, None


class BinaryQuantizer(autograd.Function):

    @staticmethod
    def forward(ctx, input, num_bits, min_value=None, max_value=None, num_groups=1):
        assert min_value is None and max_value is None
        x = input.reshape(num_groups, -1)
        m = x.norm(p=1, dim=1, keepdim=True).div(x.shape[1])
        return x.sign().mul(m).reshape(input.shape).contiguous()

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output.clone(), None, None, None, None
