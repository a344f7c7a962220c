"""Several :class:`ZeroShardedOptimizer` instances behind one optimizer facade.

Needed when parameters of one model live in different reduction domains -- the dense parameters are
sharded / averaged over the data-parallel group while MoE expert parameters are sharded / averaged over
their *expert-data-parallel* group (reference: separate ``moe`` param groups inside
``stage_1_and_2.py:304-425`` and ``engine.py:1240-1263``).  Gradient clipping uses ONE global norm: every
instance reports its squared norm, expert domains are additionally summed over the expert-parallel group,
and all instances apply the same coefficient.
"""
from typing import List

import torch

from deepspeed_b200 import comm as dist


class ZeroOptimizerGroup:

    def __init__(self, parts: List, ep_groups=None):
        self.parts = parts
        self.ep_groups = ep_groups or [None] * len(parts)
        self.custom_loss_scaler = False

    # ---- forwarding helpers ---------------------------------------------------------------------------
    @property
    def param_groups(self):
        out = []
        for p in self.parts:
            out.extend(p.param_groups)
        return out

    @property
    def loss_scale(self):
        return self.parts[0].loss_scale

    cur_scale = loss_scale

    @property
    def loss_scaler(self):
        return self.parts[0].loss_scaler

    @property
    def overflow(self):
        return any(p.overflow for p in self.parts)

    @property
    def fused_in_backward(self):
        return all(p.fused_in_backward for p in self.parts)

    @property
    def _symm(self):
        return self.parts[0]._symm

    @property
    def stage(self):
        return self.parts[0].stage

    def __getattr__(self, name):
        # anything else (units, master, ...) refers to the dense domain
        return getattr(self.parts[0], name)

    def backward(self, loss, retain_graph=False):
        for p in self.parts:
            p._in_backward = True
        self.parts[0].loss_scaler.backward(loss.float(), retain_graph=retain_graph)
        for p in self.parts:
            p.end_backward()

    def zero_grad(self, set_to_none=True):
        for p in self.parts:
            p.zero_grad(set_to_none)

    def set_no_sync(self, on):
        for p in self.parts:
            p.set_no_sync(on)

    def set_gradient_accumulation_steps(self, gas):
        for p in self.parts:
            p.set_gradient_accumulation_steps(gas)

    def set_forced_boundary(self, is_boundary):
        for p in self.parts:
            p.set_forced_boundary(is_boundary)

    def disable_fused_in_backward(self, reason):
        for p in self.parts:
            p.disable_fused_in_backward(reason)

    def step(self, closure=None):
        for p in self.parts:
            p.prepare_step()
        if any(p.needs_norm() for p in self.parts):
            dev = self.parts[0].stats.sumsq.device
            total = torch.zeros(1, dtype=torch.float32, device=dev)
            inf = torch.zeros(1, dtype=torch.int32, device=dev)
            contribs = []
            for p, epg in zip(self.parts, self.ep_groups):
                ss, fi = p.stats.sumsq.clone().to(dev), p.stats.found_inf.clone().to(dev)
                if epg is not None and dist.get_world_size(epg) > 1:
                    dist.all_reduce(ss, group=epg)  # experts differ across the EP group: sum their norms
                    dist.all_reduce(fi, op=dist.ReduceOp.MAX, group=epg)
                contribs.append((ss, fi))
                total += ss
                inf = torch.maximum(inf, fi)
            for p, (ss, fi) in zip(self.parts, contribs):
                p.stats.sumsq.zero_()  # finish_step adds `extra` to the instance's own value
                p.finish_step(extra_sumsq=total, extra_found_inf=inf)
        else:
            for p in self.parts:
                p.finish_step()

    def get_global_grad_norm(self):
        return self.parts[0].get_global_grad_norm()

    def state_dict(self, layout=None):
        # several reduction domains: always the arena layout (the reference's MoE shards are organised per expert group)
        return {"multi": [p.state_dict(layout="arena") for p in self.parts], "names": [p.name for p in self.parts]}

    def load_state_dict(self, sd, load_optimizer_states=True, load_from_fp32_weights=True, param_shapes=None):
        for p, s in zip(self.parts, sd["multi"]):
            p.load_state_dict(s, load_optimizer_states, load_from_fp32_weights)

    def destroy(self):
        for p in self.parts:
            p.destroy()

    def gather_param_temp(self, param):
        ref = getattr(param, "_ds_zero", None)
        return ref().gather_param_temp(param)

    def release_param_temp(self, param, write_back_from=None):
        ref = getattr(param, "_ds_zero", None)
        return ref().release_param_temp(param, write_back_from)
