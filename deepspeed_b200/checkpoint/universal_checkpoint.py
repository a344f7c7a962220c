"""Load a universal checkpoint into an engine of ANY parallel shape.

Reference: ``checkpoint/universal_checkpoint.py:22 load_hp_checkpoint_state`` (per-parameter slicing for TP,
vocab padding) and the optimizers' ``_load_universal_checkpoint``.  Every rank reads only the parameters that
overlap its arena shard, slices the overlapping range, and copies it into its fp32 master / optimizer state.
"""
import os
import re
from dataclasses import dataclass

import torch

from deepspeed_b200.runtime.zero.units import param_fragments
from deepspeed_b200.utils.logging import log_dist, logger


@dataclass
class SubparamShape:
    """Shape recipe of a parameter that is the concatenation of sub-parameters with their own TP slicing (universal
    checkpoint metadata, reference ``checkpoint/universal_checkpoint.py:16``).  An entry of ``shape`` may be a tuple: the
    sizes of the fused sub-parameters along ``partition_dim``."""
    patterns: list
    shape: tuple
    partition_dim: int


def _tp_slice(t, blob, tp_rank, tp_world_size, tp_dim):
    """Cut this rank's TP slice out of a merged tensor, honouring the merge metadata ``ds_to_universal`` recorded."""
    n_sub = blob.get("param_n_sub_params")
    sub = blob.get("sub_param_shape")
    dim = blob.get("cat_dim", tp_dim) if tp_dim is None else tp_dim
    if sub is not None:
        pd = sub.partition_dim
        sizes = sub.shape[pd] if isinstance(sub.shape[pd], tuple) else (sub.shape[pd], )
        full = [sum(d) if isinstance(d, tuple) else d for d in sub.shape]
        v = t.view(full)
        cols, at = [], 0
        for sz in sizes:
            cols.append(v.narrow(pd, at, sz).chunk(tp_world_size, dim=pd)[tp_rank])
            at += sz
        return torch.cat(cols, dim=pd)
    if n_sub:
        subs = t.chunk(n_sub, dim=dim)
        return torch.cat([x.chunk(tp_world_size, dim=dim)[tp_rank] for x in subs], dim=dim)
    if dim is None:
        return t  # replicated / averaged parameter
    return t.chunk(tp_world_size, dim=dim)[tp_rank]


def load_hp_checkpoint_state(folder, key, full_shape, tp_rank=0, tp_world_size=1, tp_dim=None, vocab_pad_to=None):
    """Read ``<folder>/<key>.pt`` and return the (TP-sliced, vocab-padded) full-precision tensor."""
    blob = torch.load(os.path.join(folder, f"{key}.pt"), map_location="cpu", weights_only=False)
    meta = blob if isinstance(blob, dict) else {}
    t = blob["param"] if isinstance(blob, dict) else blob
    if vocab_pad_to is None and meta.get("vocab_tensor") and tp_world_size >= 1 and len(full_shape) >= 1:
        # the merged vocabulary was trimmed to the tokenizer size: re-pad so it divides over the TP ranks
        rows = full_shape[0] * (tp_world_size if (tp_dim in (None, 0) and meta.get("cat_dim", 0) == 0) else 1)
        vocab_pad_to = rows if rows > t.shape[0] else None
    if vocab_pad_to is not None and t.dim() >= 1 and t.shape[0] < vocab_pad_to:
        pad = torch.zeros(vocab_pad_to - t.shape[0], *t.shape[1:], dtype=t.dtype)
        t = torch.cat([t, pad], 0)
    if tp_world_size > 1 and (tp_dim is not None or meta.keys() & {"cat_dim", "param_n_sub_params", "sub_param_shape"}):
        t = _tp_slice(t, meta, tp_rank, tp_world_size, tp_dim)
    if tuple(t.shape) != tuple(full_shape):
        if t.numel() == torch.Size(full_shape).numel():
            t = t.reshape(full_shape)
        else:
            raise ValueError(f"{folder}: checkpoint shape {tuple(t.shape)} does not match parameter {tuple(full_shape)}")
    return t


def enable_universal_checkpoint(param_list):
    """Attach ``param.load_hp_checkpoint_state(folder, tp_rank, tp_world_size, key="fp32")`` to every parameter
    (reference ``universal_checkpoint.py:144``)."""
    import types

    def _bound(self, folder, tp_rank=0, tp_world_size=1, key="fp32"):
        return load_hp_checkpoint_state(folder, key, self.shape, tp_rank, tp_world_size)

    for p in param_list:
        p.load_hp_checkpoint_state = types.MethodType(_bound, p)


def load_universal_into_optimizer(zo, zero_dir, load_optimizer_states=True, refresh=True):
    """Copy this rank's fragments of every parameter's fp32 weight (+ optimizer moments) from the per-parameter folders of a
    universal checkpoint into the sharded optimizer's arenas.  Returns the number of parameters touched."""
    rank = zo.shard_rank
    states = zo.flat_opt.state_tensors() if hasattr(zo.flat_opt, "state_tensors") else {}
    loaded = 0
    with torch.no_grad():
        for u in zo.units:
            for s in u.slots:
                pdir = os.path.join(zero_dir, s.name)
                if not os.path.isdir(pdir):
                    logger.warning(f"universal checkpoint has no entry for {s.name}; keeping initial value")
                    continue
                frags = [f for f in param_fragments(u, s, zo.shard_world) if f[0] == rank]
                if not frags:
                    continue
                full = load_hp_checkpoint_state(pdir, "fp32", s.shape).reshape(-1)
                extra = {}
                if load_optimizer_states:
                    for k in states:
                        if os.path.isfile(os.path.join(pdir, f"{k}.pt")):
                            extra[k] = load_hp_checkpoint_state(pdir, k, s.shape).reshape(-1)
                for (_, pstart, astart, length) in frags:
                    piece = full[pstart:pstart + length]
                    if zo.master is not None:
                        zo.master[astart:astart + length].copy_(piece)
                    else:
                        sh = zo._lp_shard(u)
                        o = astart - u.arena_offset
                        sh[o:o + length].copy_(piece.to(sh.dtype))
                    for k, t in extra.items():
                        states[k][astart:astart + length].copy_(t[pstart:pstart + length])
                loaded += 1
        if refresh:
            zo._refresh_lp_from_master()
    return loaded


def load_universal_into_engine(engine, load_dir, tag, load_optimizer_states=True):
    folder = os.path.join(load_dir, str(tag))
    if not os.path.isdir(os.path.join(folder, "zero")):
        lu = os.path.join(load_dir, "latest_universal")
        if os.path.isfile(lu):
            with open(lu) as f:
                folder = os.path.join(load_dir, f.read().strip())
    zero_dir = os.path.join(folder, "zero")
    if not os.path.isdir(zero_dir):
        raise FileNotFoundError(f"{folder} is not a universal checkpoint (no zero/ directory); run ds_to_universal")
    ms = torch.load(os.path.join(folder, "mp_rank_00_model_states.pt"), map_location="cpu", weights_only=False)
    zo = engine.optimizer
    # buffers + scheduler + counters
    bufs = {k: v for k, v in ms["module"].items() if k in set(ms.get("buffer_names", []))}
    if bufs:
        engine.module.load_state_dict(bufs, strict=False)
    engine.global_steps = ms.get("global_steps", 0)
    engine.global_samples = ms.get("global_samples", 0)
    engine.skipped_steps = ms.get("skipped_steps", 0)
    if engine.lr_scheduler is not None and ms.get("lr_scheduler") is not None:
        engine.lr_scheduler.load_state_dict(ms["lr_scheduler"])
    if zo is None or not hasattr(zo, "units"):
        sd = {}
        for name in os.listdir(zero_dir):
            sd[name] = torch.load(os.path.join(zero_dir, name, "fp32.pt"), map_location="cpu", weights_only=False)["param"]
        engine.module.load_state_dict(sd, strict=False)
        return folder, {}
    loaded = load_universal_into_optimizer(zo, zero_dir, load_optimizer_states, refresh=False)
    rank = zo.shard_rank
    with torch.no_grad():
        meta = ms.get("optimizer_meta") or {}
        if load_optimizer_states and meta:
            if meta.get("group_steps") is not None:
                zo.group_steps = list(meta["group_steps"])[:len(zo.group_steps)] + zo.group_steps[len(meta["group_steps"]):]
            zo.global_step = meta.get("global_step", zo.global_step)
            if meta.get("loss_scaler") is not None:
                zo.loss_scaler.load_state_dict(meta["loss_scaler"])
            for g, saved in zip(zo.param_groups, meta.get("param_groups") or []):
                g.update({k: v for k, v in saved.items() if k != "params"})
        zo._refresh_lp_from_master()
    log_dist(f"loaded universal checkpoint {folder}: {loaded} parameters touched on rank {rank}", ranks=[0])
    reserved = {"module", "buffer_names", "optimizer", "param_shapes", "frozen_param_shapes", "frozen_param_fragments",
                "shared_params", "lr_scheduler", "data_sampler", "random_ltd", "sparse_tensor_module_names", "skipped_steps",
                "global_steps", "global_samples", "dp_world_size", "mp_world_size", "ds_config", "ds_version", "ds_b200_layout",
                "num_experts", "universal_checkpoint_info", "optimizer_meta"}
    return folder, {k: v for k, v in ms.items() if k not in reserved}
