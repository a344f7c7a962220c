"""Host Adagrad (reference: ``ops/adagrad/cpu_adagrad.py`` + ``csrc/adagrad/cpu_adagrad.cpp``, N3)."""
import torch

from deepspeed_b200.ops import native as N
from deepspeed_b200.ops.kernels import flat_ops


def cpu_adagrad_flat(p, g, h, out=None, *, lr, eps, weight_decay, grad_scale=1.0):
    try:
        lib = N.cpu()
    except Exception:
        lib = None
    if lib is None or p.dtype != torch.float32:
        flat_ops.adagrad_flat(p, g, h, out, lr=lr, eps=eps, weight_decay=weight_decay, grad_scale=grad_scale)
        return
    rc = lib.dsb_cpu_adagrad(N.ptr(p), N.ptr(g), N.ptr(h), N.ptr(out), N.c_i64(p.numel()), N.dt(g),
                             N.dt(out) if out is not None else N.BF16, N.c_f(lr), N.c_f(eps), N.c_f(weight_decay),
                             N.c_f(grad_scale))
    if rc != 0:
        raise RuntimeError(f"dsb_cpu_adagrad failed: {rc}")


class DeepSpeedCPUAdagrad(torch.optim.Optimizer):
    optimizer_id = 0

    def __init__(self, model_params, lr=1e-2, eps=1e-10, weight_decay=0, amsgrad=False, fp32_optimizer_states=True):
        super().__init__(model_params, dict(lr=lr, eps=eps, weight_decay=weight_decay, amsgrad=amsgrad))
        self.opt_id = DeepSpeedCPUAdagrad.optimizer_id
        DeepSpeedCPUAdagrad.optimizer_id += 1

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = 0
                    st["exp_avg_sq"] = torch.zeros_like(p, dtype=torch.float32)
                st["step"] += 1
                if p.grad.is_sparse:
                    g = p.grad.coalesce()
                    idx = g.indices()[0]
                    rows = p.data[idx].float()
                    gv = g.values().float()
                    hv = st["exp_avg_sq"][idx]
                    hv.add_(gv * gv)
                    rows.addcdiv_(gv, hv.sqrt() + group["eps"], value=-group["lr"])
                    st["exp_avg_sq"][idx] = hv
                    p.data[idx] = rows.to(p.dtype)
                    continue
                cpu_adagrad_flat(p.data.view(-1), p.grad.contiguous().view(-1), st["exp_avg_sq"].view(-1), None,
                                 lr=group["lr"], eps=group["eps"], weight_decay=group["weight_decay"])
        return loss
