"""``FusedLamb`` (reference: ``ops/lamb/fused_lamb.py`` + ``csrc/lamb/fused_lamb_cuda_kernel.cu``, N4).

Two launches per tensor: phase 1 computes the Adam direction and per-block partial norms, phase 2
reduces the partials (every block, deterministic) into the trust ratio and applies the update."""
import torch

from deepspeed_b200.ops.kernels import flat_ops


class FusedLamb(torch.optim.Optimizer):

    def __init__(self, params, lr=1e-3, bias_correction=True, betas=(0.9, 0.999), eps=1e-8, eps_inside_sqrt=False,
                 weight_decay=0.0, max_grad_norm=0.0, max_coeff=10.0, min_coeff=0.01, amsgrad=False):
        if amsgrad:
            raise RuntimeError("FusedLamb does not support the AMSGrad variant.")
        defaults = dict(lr=lr, bias_correction=bias_correction, betas=betas, eps=eps, weight_decay=weight_decay,
                        max_grad_norm=max_grad_norm, max_coeff=max_coeff, min_coeff=min_coeff)
        super().__init__(params, defaults)
        self.eps_mode = 0 if eps_inside_sqrt else 1
        self.lamb_coeffs = []

    @torch.no_grad()
    def step(self, closure=None, grads=None, output_params=None, scale=1.0, grad_norms=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        self.lamb_coeffs = []
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, dtype=torch.float32)
                    st["exp_avg_sq"] = torch.zeros_like(p, dtype=torch.float32)
                st["step"] += 1
                gs = 1.0 / scale
                if group["max_grad_norm"] > 0:
                    gn = float(p.grad.float().norm()) / scale
                    clip = gn / group["max_grad_norm"]
                    if clip > 1:
                        gs = gs / clip
                c = flat_ops.lamb_flat(p.data.view(-1), p.grad.contiguous().view(-1), st["exp_avg"].view(-1),
                                       st["exp_avg_sq"].view(-1), None, lr=group["lr"], beta1=b1, beta2=b2,
                                       eps=group["eps"], weight_decay=group["weight_decay"], step=st["step"],
                                       bias_correction=group["bias_correction"], max_coeff=group["max_coeff"],
                                       min_coeff=group["min_coeff"], grad_scale=gs)
                self.lamb_coeffs.append(c)
        return loss

    def get_lamb_coeffs(self):
        return [float(c.item()) for c in self.lamb_coeffs]
