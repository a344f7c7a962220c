"""1-bit LAMB (reference ``runtime/fp16/onebit/lamb.py:15 OnebitLamb``).

Warm-up: LAMB (Adam direction scaled per tensor by ``clamp(||w|| / ||update||)``), recording an exponential
average of each tensor's trust ratio.  Compression stage: variance and the *base* trust ratios are frozen;
momentum is exchanged 1-bit after being pre-scaled so all tensors share a similar magnitude (``scaling_coeff``),
and the live trust ratio is the frozen one modulated by ``factor`` — the ratio between the frozen variance and a
locally refreshed variance estimate — clamped to ``[factor_min, factor_max]`` and rate-limited by
``factor_threshold``.
"""
import torch

from ._base import _CompressedOptimizer


class OnebitLamb(_CompressedOptimizer):

    def __init__(self, params, deepspeed=None, lr=1e-3, freeze_step=100000, bias_correction=True, betas=(0.9, 0.999),
                 eps=1e-8, eps_inside_sqrt=False, weight_decay=0.0, max_grad_norm=0.0, max_coeff=10.0, min_coeff=0.01,
                 amsgrad=False, cuda_aware=False, comm_backend_name="nccl", coeff_beta=0.9, factor_max=4.0, factor_min=0.5,
                 factor_threshold=0.1):
        if amsgrad:
            raise RuntimeError("1-bit Lamb does not support the AMSGrad variant.")
        defaults = dict(lr=lr, bias_correction=bias_correction, betas=betas, eps=eps, weight_decay=weight_decay,
                        max_grad_norm=max_grad_norm, max_coeff=max_coeff, min_coeff=min_coeff)
        super().__init__(params, defaults)
        self.eps_mode = 0 if eps_inside_sqrt else 1
        self.freeze_step = freeze_step
        self.lamb_freeze_key = False
        self.initialize = False
        self.coeff_beta = coeff_beta
        self.factor_max, self.factor_min, self.factor_threshold = factor_max, factor_min, factor_threshold
        self.lamb_coeffs = []
        self._setup(deepspeed, cuda_aware, comm_backend_name)

    @property
    def freeze_key(self):
        return self.lamb_freeze_key

    def get_lamb_coeffs(self):
        return self.lamb_coeffs

    @torch.no_grad()
    def step(self, closure=None, grads=None):
        loss = closure() if closure is not None else None
        self.lamb_coeffs = []
        step_now = 0
        entering = False
        for group in self.param_groups:
            b1, b2 = group["betas"]
            lr, eps, wd = group["lr"], group["eps"], group["weight_decay"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                g = p.grad.float()
                st = self.state[p]
                if not st:
                    st.update(step=0, lamb_coeff_freeze=0.0, last_factor=1.0,
                              exp_avg=torch.zeros_like(p, dtype=torch.float32),
                              exp_avg_sq=torch.zeros_like(p, dtype=torch.float32),
                              exp_avg_sq_fresh=torch.zeros_like(p, dtype=torch.float32))
                st["step"] += 1
                step_now = st["step"]
                m, v, vf = st["exp_avg"], st["exp_avg_sq"], st["exp_avg_sq_fresh"]
                w32 = p.float()
                if not self.lamb_freeze_key:
                    m.mul_(b1).add_(g, alpha=1 - b1)
                    v.mul_(b2).addcmul_(g, g, value=1 - b2)
                    if st["step"] == self.freeze_step:
                        vf.copy_(v)
                    denom = (v + eps).sqrt() if self.eps_mode == 0 else v.sqrt().add_(eps)
                    upd = m / denom
                    if wd > 0.0:
                        upd = upd + wd * w32
                    wn, un = w32.norm(), upd.norm()
                    coeff = 1.0
                    if wn != 0 and un != 0:
                        coeff = float((wn / un).clamp(group["min_coeff"], group["max_coeff"]))
                        if st["step"] == 1:
                            st["lamb_coeff_freeze"] = coeff
                        else:
                            st["lamb_coeff_freeze"] = self.coeff_beta * st["lamb_coeff_freeze"] + (1 - self.coeff_beta) * coeff
                    self.lamb_coeffs.append(coeff)
                    p.add_(upd.to(p.dtype), alpha=-lr * coeff)
                else:
                    if "scaling_coeff" not in st:
                        entering = True
                        st["scaling_coeff"] = 1.0
                    prev = m.clone()
                    m.mul_(b1).add_(g, alpha=1 - b1)
                    scaled = m * st["scaling_coeff"]
                    avg = self._compressed_mean(scaled, st, p) / st["scaling_coeff"]
                    if "exp_avg_mask" in group:
                        avg = avg * group["exp_avg_mask"].to(avg.device)
                    # refresh a local variance estimate from the momentum delta (gradient reconstruction)
                    g_rec = (avg - b1 * prev) / (1 - b1)
                    m.copy_(avg)
                    vf.mul_(b2).addcmul_(g_rec, g_rec, value=1 - b2)
                    denom = v.sqrt().add_(eps)
                    denom_fresh = vf.sqrt().add_(eps)
                    factor = float((denom / denom_fresh).max())
                    factor = min(max(factor, self.factor_min), self.factor_max)
                    lo, hi = st["last_factor"] * (1 - self.factor_threshold), st["last_factor"] * (1 + self.factor_threshold)
                    factor = min(max(factor, lo), hi)
                    st["last_factor"] = factor
                    coeff = st["lamb_coeff_freeze"] * factor
                    self.lamb_coeffs.append(coeff)
                    upd = m / denom
                    if wd > 0.0:
                        upd = upd + wd * w32
                    p.add_(upd.to(p.dtype), alpha=-lr * coeff)
        if entering:
            # momentum pre-scaling: bring every tensor's RMS to the global average so one sign-scale fits all
            rms = {id(p): float(self.state[p]["exp_avg"].norm() / max(1, p.numel())**0.5)
                   for g in self.param_groups for p in g["params"] if p in self.state and "exp_avg" in self.state[p]}
            if rms:
                united = sum(rms.values()) / len(rms)
                for g in self.param_groups:
                    for p in g["params"]:
                        if id(p) in rms and rms[id(p)] > 0:
                            self.state[p]["scaling_coeff"] = united / rms[id(p)]
        if not self.lamb_freeze_key and step_now >= self.freeze_step:
            self.lamb_freeze_key = True
            self._set_engine_allreduce(False)
        self.initialize = True
        return loss

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        for st in self.state.values():
            for k in ("worker_error", "server_error"):
                st.pop(k, None)
        any_state = next(iter(self.state.values()), None)
        self.lamb_freeze_key = bool(any_state is not None and any_state.get("step", 0) >= self.freeze_step)
        self._set_engine_allreduce(not self.lamb_freeze_key)
