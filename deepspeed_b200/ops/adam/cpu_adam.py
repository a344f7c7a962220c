"""Host Adam (``DeepSpeedCPUAdam``) backed by the AVX-512/AVX2 + OpenMP kernel in
``csrc/cpu/cpu_optim.cpp``.

Parity target: reference ``ops/adam/cpu_adam.py`` + ``csrc/adam/cpu_adam_impl.cpp`` (N2).  The
native entry point works on flat host buffers (fp32 master / states, fp32|bf16|fp16 gradients) and
can emit the bf16/fp16 copy of the updated parameters in the same pass, which is what the ZeRO
offload tier H2D-copies back to the GPU.
"""
import ctypes

import torch

from deepspeed_b200.ops import native as N
from deepspeed_b200.ops.kernels import flat_ops

_lib = None
_tried = False


def _load():
    global _lib, _tried
    if _tried:
        return _lib
    _tried = True
    try:
        _lib = N.cpu()
        _lib.dsb_cpu_adam.restype = ctypes.c_int
    except Exception:
        _lib = None
    return _lib


def available() -> bool:
    return _load() is not None


_threads_set = False


def configure_threads(n=None):
    """Give this rank's host optimizer its share of the cores.  ``torchrun`` exports ``OMP_NUM_THREADS=1`` for every rank;
    honouring that would run the offload tier's Adam on one core.  ``DSB200_CPU_THREADS`` overrides; default =
    usable cores // local world size (the reference's launcher does the same core split, ``launcher/launch.py:227``)."""
    global _threads_set
    lib = _load()
    if lib is None:
        return 0
    if n is None:
        import os
        env = os.environ.get("DSB200_CPU_THREADS")
        if env:
            n = int(env)
        else:
            try:
                cores = len(os.sched_getaffinity(0))
            except AttributeError:
                cores = os.cpu_count() or 1
            local = int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("LOCAL_SIZE", "1")) or 1)
            n = max(1, min(64, cores // max(1, local)))
    lib.dsb_cpu_set_threads(int(n))
    _threads_set = True
    return int(n)


def cpu_adam_flat(p, g, m, v, out=None, *, lr, beta1, beta2, eps, weight_decay, step, adamw=True, bias_correction=True,
                  grad_scale=1.0, d_gscale=None, d_skip=None):
    """Adam on flat host tensors (``p``/``m``/``v`` fp32)."""
    if d_skip is not None and int(d_skip.item()) != 0:
        return
    gs = grad_scale * (float(d_gscale.item()) if d_gscale is not None else 1.0)
    lib = _load()
    ok = (lib is not None and p.dtype == torch.float32 and m.dtype == torch.float32 and p.is_contiguous()
          and g.is_contiguous())
    if not ok:
        flat_ops.adam_flat(p, g, m, v, out, lr=lr, beta1=beta1, beta2=beta2, eps=eps, weight_decay=weight_decay,
                           step=step, adamw=adamw, bias_correction=bias_correction, grad_scale=gs)
        return
    bc1 = 1.0 - beta1**step if bias_correction else 1.0
    bc2 = 1.0 - beta2**step if bias_correction else 1.0
    rc = lib.dsb_cpu_adam(N.ptr(p), N.ptr(g), N.ptr(m), N.ptr(v), N.ptr(out), N.c_i64(p.numel()), N.dt(g),
                          N.dt(out) if out is not None else N.BF16, N.c_f(lr), N.c_f(beta1), N.c_f(beta2), N.c_f(eps),
                          N.c_f(weight_decay), N.c_f(bc1), N.c_f(bc2), int(bool(adamw)), N.c_f(gs))
    if rc != 0:
        raise RuntimeError(f"dsb_cpu_adam failed with code {rc}")


class DeepSpeedCPUAdam(torch.optim.Optimizer):
    optimizer_id = 0

    def __init__(self, model_params, lr=1e-3, bias_correction=True, betas=(0.9, 0.999), eps=1e-8, weight_decay=0,
                 amsgrad=False, adamw_mode=True, fp32_optimizer_states=True):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, bias_correction=bias_correction,
                        amsgrad=amsgrad)
        super().__init__(model_params, defaults)
        self.opt_id = DeepSpeedCPUAdam.optimizer_id
        DeepSpeedCPUAdam.optimizer_id += 1
        self.adam_w_mode = adamw_mode
        self.fp32_optimizer_states = fp32_optimizer_states

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
                assert p.device.type == "cpu", "CPUAdam param is on a non-cpu device; use FusedAdam for GPU params"
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = 0
                    sdt = torch.float32 if self.fp32_optimizer_states else p.dtype
                    st["exp_avg"] = torch.zeros_like(p, dtype=sdt)
                    st["exp_avg_sq"] = torch.zeros_like(p, dtype=sdt)
                st["step"] += 1
                b1, b2 = group["betas"]
                if p.dtype == torch.float32:
                    cpu_adam_flat(p.data.view(-1), p.grad.data.contiguous().view(-1), st["exp_avg"].view(-1),
                                  st["exp_avg_sq"].view(-1), None, lr=group["lr"], beta1=b1, beta2=b2,
                                  eps=group["eps"], weight_decay=group["weight_decay"], step=st["step"],
                                  adamw=self.adam_w_mode, bias_correction=group["bias_correction"])
                else:
                    flat_ops.adam_flat(p.data.view(-1), p.grad.data.contiguous().view(-1), st["exp_avg"].view(-1),
                                       st["exp_avg_sq"].view(-1), None, lr=group["lr"], beta1=b1, beta2=b2,
                                       eps=group["eps"], weight_decay=group["weight_decay"], step=st["step"],
                                       adamw=self.adam_w_mode, bias_correction=group["bias_correction"])
        return loss


# ---- binding-level API of the reference op (``csrc/adam/cpu_adam.cpp``: create_adam / adam_update / destroy_adam) ----------
_registry = {}


def create_adam(optimizer_id, alpha=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0, adamw_mode=True,
                should_log=False):
    _registry[optimizer_id] = dict(lr=alpha, beta1=beta1, beta2=beta2, eps=eps, weight_decay=weight_decay, adamw=adamw_mode)
    return 0


def adam_update(optimizer_id, step, lr, beta1, beta2, epsilon, weight_decay, bias_correction, params, grads, exp_avg,
                exp_avg_sq):
    cfg = _registry[optimizer_id]
    cpu_adam_flat(params.view(-1), grads.view(-1), exp_avg.view(-1), exp_avg_sq.view(-1), None, lr=lr, beta1=beta1,
                  beta2=beta2, eps=epsilon, weight_decay=weight_decay, step=step, adamw=cfg["adamw"],
                  bias_correction=bool(bias_correction))
    return 0


def adam_update_copy(optimizer_id, step, lr, beta1, beta2, epsilon, weight_decay, bias_correction, params, grads, exp_avg,
                     exp_avg_sq, device_params):
    cfg = _registry[optimizer_id]
    lp = torch.empty(params.numel(), dtype=device_params.dtype)
    cpu_adam_flat(params.view(-1), grads.view(-1), exp_avg.view(-1), exp_avg_sq.view(-1), lp, lr=lr, beta1=beta1, beta2=beta2,
                  eps=epsilon, weight_decay=weight_decay, step=step, adamw=cfg["adamw"], bias_correction=bool(bias_correction))
    device_params.copy_(lp.view_as(device_params), non_blocking=True)
    return 0


def destroy_adam(optimizer_id):
    _registry.pop(optimizer_id, None)
    return 0
