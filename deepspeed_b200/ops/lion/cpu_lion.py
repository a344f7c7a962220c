"""Host Lion (reference: ``ops/lion/cpu_lion.py`` + ``csrc/lion/cpu_lion_impl.cpp``, N3)."""
import ctypes

import torch

from deepspeed_b200.ops import native as N
from deepspeed_b200.ops.kernels import flat_ops


def _native():
    try:
        return N.cpu()
    except Exception:
        return None


def cpu_lion_flat(p, g, m, out=None, *, lr, beta1, beta2, weight_decay, grad_scale=1.0):
    lib = _native()
    if lib is None or p.dtype != torch.float32 or m.dtype != torch.float32:
        flat_ops.lion_flat(p, g, m, out, lr=lr, beta1=beta1, beta2=beta2, weight_decay=weight_decay,
                           grad_scale=grad_scale)
        return
    rc = lib.dsb_cpu_lion(N.ptr(p), N.ptr(g), N.ptr(m), N.ptr(out), N.c_i64(p.numel()), N.dt(g),
                          N.dt(out) if out is not None else N.BF16, N.c_f(lr), N.c_f(beta1), N.c_f(beta2),
                          N.c_f(weight_decay), N.c_f(grad_scale))
    if rc != 0:
        raise RuntimeError(f"dsb_cpu_lion failed: {rc}")


class DeepSpeedCPULion(torch.optim.Optimizer):
    optimizer_id = 0

    def __init__(self, model_params, lr=1e-3, betas=(0.9, 0.999), weight_decay=0, fp32_optimizer_states=True):
        super().__init__(model_params, dict(lr=lr, betas=betas, weight_decay=weight_decay))
        self.opt_id = DeepSpeedCPULion.optimizer_id
        DeepSpeedCPULion.optimizer_id += 1

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
                    st["exp_avg"] = torch.zeros_like(p, dtype=torch.float32)
                if p.dtype == torch.float32:
                    cpu_lion_flat(p.data.view(-1), p.grad.contiguous().view(-1), st["exp_avg"].view(-1), None,
                                  lr=group["lr"], beta1=b1, beta2=b2, weight_decay=group["weight_decay"])
                else:
                    flat_ops.lion_flat(p.data.view(-1), p.grad.contiguous().view(-1), st["exp_avg"].view(-1), None,
                                       lr=group["lr"], beta1=b1, beta2=b2, weight_decay=group["weight_decay"])
        return loss
