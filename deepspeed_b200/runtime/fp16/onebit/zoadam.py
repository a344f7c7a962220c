"""0/1 Adam (reference ``runtime/fp16/onebit/zoadam.py:14 ZeroOneAdam``).

Two knobs on top of 1-bit Adam: (1) *adaptive variance freezing* — the variance is refreshed with a dense
all-reduced gradient only every ``var_interval`` steps (interval doubles every ``var_update_scaler`` refreshes)
and in between the gradient itself travels 1-bit; after ``var_freeze_step`` it is frozen for good.  (2) *local
steps* — after the freeze, ranks run ``local_step_interval`` purely local steps (interval doubles every
``local_step_scaler`` steps up to ``local_step_clipper``) and synchronise the accumulated update with one 1-bit
all-reduce.
"""
import torch

from ._base import _CompressedOptimizer


class ZeroOneAdam(_CompressedOptimizer):

    def __init__(self, params, deepspeed=None, lr=1e-3, bias_correction=True, betas=(0.9, 0.999), eps=1e-8,
                 eps_inside_sqrt=False, weight_decay=0.0, max_grad_norm=0.0, var_freeze_step=100000, var_update_scaler=16,
                 local_step_scaler=32678, local_step_clipper=16, amsgrad=False, cuda_aware=False, comm_backend_name="nccl"):
        if amsgrad:
            raise RuntimeError("0/1 Adam does not support the AMSGrad variant.")
        defaults = dict(lr=lr, bias_correction=bias_correction, betas=betas, eps=eps, weight_decay=weight_decay,
                        max_grad_norm=max_grad_norm)
        super().__init__(params, defaults)
        self.eps_mode = 0 if eps_inside_sqrt else 1
        self.var_freeze_step = var_freeze_step
        self.var_update_scaler = var_update_scaler
        self.local_step_scaler = local_step_scaler
        self.local_step_clipper = local_step_clipper
        self.freeze_key = False
        self.reinitial_error_buffer = False
        self.initialize = False
        self._setup(deepspeed, cuda_aware, comm_backend_name)

    @torch.no_grad()
    def step(self, closure=None, grads=None):
        loss = closure() if closure is not None else None
        last_step, next_dense = 0, True
        for group in self.param_groups:
            b1, b2 = group["betas"]
            lr, eps, wd = group["lr"], group["eps"], group["weight_decay"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                g = p.grad.float()
                st = self.state[p]
                if not st:
                    st.update(step=0, exp_avg=torch.zeros_like(p, dtype=torch.float32),
                              exp_avg_sq=torch.zeros_like(p, dtype=torch.float32), var_interval=1, var_counter=0,
                              local_step_interval=1, local_step_counter=0, lrs=0.0,
                              momentum_accumulator=torch.zeros_like(p, dtype=torch.float32))
                st["step"] += 1
                last_step = st["step"]
                m, v, acc = st["exp_avg"], st["exp_avg_sq"], st["momentum_accumulator"]
                if not self.freeze_key:
                    if st["step"] % st["var_interval"] == 0:
                        # dense step: the engine all-reduced this gradient
                        v.mul_(b2).addcmul_(g, g, value=1 - b2)
                        m.mul_(b1).add_(g, alpha=1 - b1)
                    else:
                        g1 = self._compressed_mean(g, st, p)
                        if "exp_avg_mask" in group:
                            g1 = g1 * group["exp_avg_mask"].to(g1.device)
                        m.mul_(b1).add_(g1, alpha=1 - b1)
                else:
                    m.mul_(b1).add_(g, alpha=1 - b1)
                    st["lrs"] += lr
                denom = (v + eps).sqrt() if self.eps_mode == 0 else v.sqrt().add_(eps)
                upd = m / denom
                if wd > 0.0:
                    upd = upd + wd * p.float()
                p.add_(upd.to(p.dtype), alpha=-lr)
                if self.freeze_key:
                    acc.add_(upd, alpha=-lr)
                    if st["step"] % st["local_step_interval"] == 0:
                        # undo the local drift, average the accumulated update (in momentum units), re-apply
                        p.sub_(acc.to(p.dtype))
                        acc.mul_(denom)
                        synced = self._compressed_mean(acc, st, p)
                        if "exp_avg_mask" in group:
                            synced = synced * group["exp_avg_mask"].to(synced.device)
                        m.copy_(synced).div_(-st["lrs"])
                        p.add_((synced / denom).to(p.dtype))
                        acc.zero_()
                        st["lrs"] = 0.0
                # ---- schedule bookkeeping
                if not self.freeze_key:
                    if st["step"] % st["var_interval"] == 0:
                        st["var_counter"] += 1
                        if st["var_counter"] == self.var_update_scaler:
                            st["var_counter"] = 0
                            st["var_interval"] *= 2
                    next_dense = (st["step"] + 1) % st["var_interval"] == 0
                else:
                    st["local_step_counter"] += 1
                    if st["local_step_counter"] == self.local_step_scaler:
                        st["local_step_counter"] = 0
                        st["local_step_interval"] = min(self.local_step_clipper, st["local_step_interval"] * 2)
        if not self.freeze_key:
            self._set_engine_allreduce(next_dense)
            if last_step > self.var_freeze_step:
                self.freeze_key = True
                self._set_engine_allreduce(False)
        if self.freeze_key and not self.reinitial_error_buffer:
            # error buffers tracked gradients so far; from now on they track accumulated momentum
            for st in self.state.values():
                if "worker_error" in st:
                    st["worker_error"].zero_()
                    st["server_error"].zero_()
            self.reinitial_error_buffer = True
        self.initialize = True
        return loss

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        for st in self.state.values():
            st.pop("worker_error", None)
            st.pop("server_error", None)
        any_state = next(iter(self.state.values()), None)
        self.freeze_key = bool(any_state is not None and any_state.get("step", 0) > self.var_freeze_step)
        self.reinitial_error_buffer = False
        self._set_engine_allreduce(not self.freeze_key)
