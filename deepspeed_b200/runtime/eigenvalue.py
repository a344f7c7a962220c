"""Dominant Hessian eigenvalue per layer by power iteration on Hessian-vector products (reference
``runtime/eigenvalue.py:13``); feeds the MoQ quantisation-period schedule."""
import torch

from deepspeed_b200.utils.logging import log_dist


class Eigenvalue:

    def __init__(self, verbose=False, max_iter=100, tol=1e-2, stability=0, gas_boundary_resolution=1, layer_name="",
                 layer_num=0):
        self.verbose, self.max_iter, self.tol, self.stability = verbose, max_iter, tol, stability
        self.gas_boundary_resolution = gas_boundary_resolution
        self.layer_name, self.layer_num = layer_name, layer_num
        assert len(layer_name) > 0 and layer_num > 0
        log_dist(f"enabled eigenvalue with verbose={verbose}, max_iter={max_iter}, tol={tol}, stability={stability}, "
                 f"gas_boundary_resolution={gas_boundary_resolution}, layer_name={layer_name}, layer_num={layer_num}",
                 ranks=[0])

    @staticmethod
    def nan_to_num(x):
        return torch.nan_to_num(x, nan=0.0, posinf=0.0, neginf=0.0)

    def normalize(self, v):
        norm = torch.sqrt(sum((x * x).sum() for x in v))
        return [self.nan_to_num(x / (norm + self.stability)) for x in v]

    def inner_product(self, xs, ys):
        return sum((x * y).sum() for x, y in zip(xs, ys))

    def get_layers(self, module):
        scope = module
        for name in self.layer_name.split("."):
            scope = getattr(scope, name)
        return scope

    def compute_eigenvalue(self, module, device=None, scale=1.0):
        """Returns ``{param_id: (eigenvalue_normalised_to_max, layer_id)}`` for every parameter of the layers
        under ``layer_name``; gradients must have been produced with ``create_graph=True``."""
        block_eigen = []
        layers = self.get_layers(module)
        for li in range(self.layer_num):
            params = [p for p in layers[li].parameters() if p.grad is not None and p.grad.grad_fn is not None]
            if not params:
                block_eigen.append(0.0)
                continue
            grads = [p.grad for p in params]
            v = self.normalize([torch.randn_like(p) for p in params])
            cur, prev = 1.0, 0.0
            it = 0
            while it < self.max_iter and abs(cur) > 0 and abs((cur - prev) / cur) >= self.tol:
                prev = cur
                Hv = torch.autograd.grad(grads, params, grad_outputs=v, only_inputs=True, retain_graph=True)
                Hv = [self.nan_to_num(h.float()) for h in Hv]
                cur = float(self.inner_product(Hv, v))
                v = self.normalize(Hv)
                v = [x / scale for x in v]
                it += 1
            block_eigen.append(cur * scale)
            if self.verbose:
                log_dist(f"block: {li}, power iteration: {it}, eigenvalue: {block_eigen[-1]}", ranks=[0])
        block_eigen = self.post_process(block_eigen)
        out = {}
        for li in range(self.layer_num):
            for p in layers[li].parameters():
                out[id(p)] = (block_eigen[li], li)
        return out

    def post_process(self, values):
        m = max((abs(v) for v in values), default=0.0)
        return [abs(v) / m if (m and v != 0.0) else 1.0 for v in values]
