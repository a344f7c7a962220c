"""``FusedLion`` (reference: ``ops/lion/fused_lion.py`` + ``csrc/lion/multi_tensor_lion.cu``, N4)."""
import torch

from deepspeed_b200.ops.kernels import flat_ops


class FusedLion(torch.optim.Optimizer):

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), weight_decay=0.0, set_grad_none=True):
        super().__init__(params, dict(lr=lr, betas=betas, weight_decay=weight_decay))
        self.set_grad_none = set_grad_none

    def zero_grad(self, set_to_none=None):
        super().zero_grad(set_to_none=self.set_grad_none if set_to_none is None else set_to_none)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if len(st) == 0:
                    st["exp_avg"] = torch.zeros_like(p)
                flat_ops.lion_flat(p.data.view(-1), p.grad.contiguous().view(-1), st["exp_avg"].view(-1), None,
                                   lr=group["lr"], beta1=b1, beta2=b2, weight_decay=group["weight_decay"])
        return loss
