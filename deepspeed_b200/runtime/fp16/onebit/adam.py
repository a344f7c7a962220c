"""1-bit Adam (reference ``runtime/fp16/onebit/adam.py:14 OnebitAdam``).

Warm-up (``step < freeze_step``): plain Adam on dense all-reduced gradients.  Compression stage: the variance is
frozen, every rank updates its momentum with its *local* gradient, and the momentum is averaged with the
error-compensated 1-bit all-reduce; the engine's dense gradient all-reduce is switched off.
"""
import torch

from ._base import _CompressedOptimizer


class OnebitAdam(_CompressedOptimizer):

    def __init__(self, params, deepspeed=None, lr=1e-3, freeze_step=100000, bias_correction=True, betas=(0.9, 0.999),
                 eps=1e-8, eps_inside_sqrt=False, weight_decay=0.0, max_grad_norm=0.0, amsgrad=False, cuda_aware=False,
                 comm_backend_name="nccl"):
        if amsgrad:
            raise RuntimeError("1-bit Adam does not support the AMSGrad variant.")
        defaults = dict(lr=lr, bias_correction=bias_correction, betas=betas, eps=eps, weight_decay=weight_decay,
                        max_grad_norm=max_grad_norm)
        super().__init__(params, defaults)
        self.eps_mode = 0 if eps_inside_sqrt else 1
        self.freeze_step = freeze_step
        self.adam_freeze_key = False
        self.initialize = False
        self._setup(deepspeed, cuda_aware, comm_backend_name)

    @property
    def freeze_key(self):
        return self.adam_freeze_key

    @torch.no_grad()
    def step(self, closure=None, grads=None):
        loss = closure() if closure is not None else None
        step_now = 0
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                g = p.grad.float()
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, dtype=torch.float32)
                    st["exp_avg_sq"] = torch.zeros_like(p, dtype=torch.float32)
                st["step"] += 1
                step_now = st["step"]
                m, v = st["exp_avg"], st["exp_avg_sq"]
                if not self.adam_freeze_key:
                    m.mul_(b1).add_(g, alpha=1 - b1)
                    v.mul_(b2).addcmul_(g, g, value=1 - b2)
                else:
                    m.mul_(b1).add_(g, alpha=1 - b1)
                    avg = self._compressed_mean(m, st, p)
                    if "exp_avg_mask" in group:
                        avg = avg * group["exp_avg_mask"].to(avg.device)
                    m.copy_(avg)
                denom = (v + group["eps"]).sqrt() if self.eps_mode == 0 else v.sqrt().add_(group["eps"])
                upd = m / denom
                if group["weight_decay"] > 0.0:
                    upd = upd + group["weight_decay"] * p.float()
                p.add_(upd.to(p.dtype), alpha=-group["lr"])
        if not self.adam_freeze_key and step_now >= self.freeze_step:
            self.adam_freeze_key = True
            self._set_engine_allreduce(False)
        self.initialize = True
        return loss

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        # error buffers are not meaningful across restarts (the reference resets them too)
        for st in self.state.values():
            st.pop("worker_error", None)
            st.pop("server_error", None)
        any_state = next(iter(self.state.values()), None)
        if any_state is not None and any_state.get("step", 0) >= self.freeze_step:
            self.adam_freeze_key = True
            self._set_engine_allreduce(False)
        else:
            self.adam_freeze_key = False
            self._set_engine_allreduce(True)
