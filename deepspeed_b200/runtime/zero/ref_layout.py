"""The reference's on-disk ZeRO optimizer-shard layout, produced from / scattered into the unit-sharded arenas.

The arenas of :class:`~deepspeed_b200.runtime.zero.sharded.ZeroShardedOptimizer` shard every *unit* flat buffer contiguously
over the DP group; the reference shards differently:

* **stage 3** (``runtime/zero/stage3.py:2548 _rigid_state_dict``): every *parameter* is cut into ``world`` pieces of
  ``ceil(numel / world)`` elements (the last one zero padded); rank ``r``'s file holds, per sub-group, the concatenation
  of its piece of every parameter -- ``fp32_flat_groups`` -- and a torch-optimizer ``optimizer_state_dict`` whose
  ``state[i]`` carries the moments laid out identically;
* **stage 1/2** (``stage_1_and_2.py:2155 state_dict``): every *param group* is flattened in parameter order, padded to a
  multiple of ``2 * world`` and cut into ``world`` equal partitions; rank ``r`` stores its partition without the trailing
  padding (``single_partition_of_fp32_groups``), the padded moments (``base_optimizer_state``), ``group_paddings`` and
  ``param_slice_mappings`` (name -> ``fragment_address(numel, start)`` inside the partition).

``export_reference_state`` re-partitions at save time (one fp32 all-gather per unit and state tensor, staged through the
host) so that the *stock* ``zero_to_fp32.py`` / ``ds_to_universal.py`` read this framework's checkpoints, and
``import_reference_state`` does the inverse (place this rank's pieces into a zeroed unit buffer, all-reduce, keep the own
shard) so a checkpoint written by stock DeepSpeed resumes in-engine at the same DP degree.
"""
from collections import OrderedDict
from typing import Dict, List

import torch

from deepspeed_b200 import comm as dist

OPTIMIZER_STATE_DICT = "optimizer_state_dict"
FP32_FLAT_GROUPS = "fp32_flat_groups"
SINGLE_PARTITION_OF_FP32_GROUPS = "single_partition_of_fp32_groups"
BASE_OPTIMIZER_STATE = "base_optimizer_state"
PARAM_SLICE_MAPPINGS = "param_slice_mappings"
GROUP_PADDINGS = "group_paddings"
PARTITION_COUNT = "partition_count"
ZERO_STAGE = "zero_stage"
LOSS_SCALER = "loss_scaler"
CLIP_GRAD = "clip_grad"
DS_VERSION = "ds_version"


def fragment_address(numel: int, start: int):
    """Same field names as the reference's ``utils/tensor_fragment.py:13`` (tools read ``.start`` / ``.numel``), but a
    stdlib object so the shard unpickles without this package (the stock ``zero_to_fp32.py`` loads the whole file)."""
    import types
    return types.SimpleNamespace(numel=int(numel), start=int(start))


def is_reference_layout(sd) -> bool:
    return isinstance(sd, dict) and (FP32_FLAT_GROUPS in sd or SINGLE_PARTITION_OF_FP32_GROUPS in sd)


def group_param_order(zo) -> List[list]:
    """Per optimizer param group: the (unit_rt, slot) pairs in the order this framework flattens them (unit order, then
    slot order).  ``param_shapes`` in the model-states file is written in exactly this order."""
    n_groups = len(zo.param_groups)
    out = [[] for _ in range(n_groups)]
    for rt in zo.rts:
        for s in rt.u.slots:
            if 0 <= s.group < n_groups:
                out[s.group].append((rt, s))
    return out


def _partitioned_numel(numel, world):
    return -(-numel // world)


def _gather_unit(zo, arena, rt):
    """fp32 copy of one whole unit flat buffer assembled from every rank's arena shard (host tensor)."""
    u = rt.u
    a = u.arena_offset
    shard = arena[a:a + u.shard_numel].detach().to(zo.device, torch.float32).contiguous()
    if zo.shard_world > 1:
        full = torch.empty(u.full_numel, dtype=torch.float32, device=zo.device)
        dist.all_gather_into_tensor(full, shard, group=zo.dp_group)
    else:
        full = shard
    return full.cpu()


def _state_arenas(zo) -> Dict[str, torch.Tensor]:
    """name -> rank-local arena for the fp32 master ("fp32") and every optimizer state tensor."""
    from deepspeed_b200.runtime.zero.flat_optimizers import TorchOptimizerAdapter
    out = OrderedDict()
    out["fp32"] = zo.master if zo.master is not None else zo._lp_arena_as_flat()
    if not isinstance(zo.flat_opt, TorchOptimizerAdapter):
        for k, v in zo.flat_opt.state_tensors().items():
            out[k] = v
    return out


def _torch_param_groups(zo):
    groups = []
    for i, g in enumerate(zo.param_groups):
        ng = {k: v for k, v in g.items() if k != "params"}
        ng["step"] = int(zo.group_steps[i]) if i < len(zo.group_steps) else 0
        ng["params"] = [i]
        groups.append(ng)
    return groups


@torch.no_grad()
def export_reference_state(zo) -> dict:
    """This rank's optimizer shard in the reference layout (collective: every DP rank must call it)."""
    from deepspeed_b200 import __version__
    world, rank = zo.shard_world, zo.shard_rank
    order = group_param_order(zo)
    arenas = _state_arenas(zo)
    stage3 = zo.stage == 3
    n_groups = len(order)
    # geometry -----------------------------------------------------------------------------------------------------
    if stage3:
        sizes = [sum(_partitioned_numel(s.numel, world) for _, s in lst) for lst in order]
    else:
        align = 2 * world
        raw = [sum(s.numel for _, s in lst) for lst in order]
        padded = [-(-n // align) * align for n in raw]
        sizes = [p // world for p in padded]
    outs = {k: [torch.zeros(sizes[g], dtype=torch.float32) for g in range(n_groups)] for k in arenas}
    # where does every parameter start inside its group's (per-rank / whole-group) flat?
    starts = {}
    for g, lst in enumerate(order):
        off = 0
        for _, s in lst:
            starts[id(s)] = off
            off += _partitioned_numel(s.numel, world) if stage3 else s.numel
    mappings = [OrderedDict() for _ in range(n_groups)]
    for rt in zo.rts:
        members = [s for s in rt.u.slots if 0 <= s.group < n_groups]
        if not members:
            continue
        for key, arena in arenas.items():
            full = _gather_unit(zo, arena, rt)
            for s in members:
                src = full[s.offset:s.offset + s.numel]
                dst = outs[key][s.group]
                if stage3:
                    pn = _partitioned_numel(s.numel, world)
                    lo, hi = rank * pn, min((rank + 1) * pn, s.numel)
                    if lo < hi:
                        dst[starts[id(s)]:starts[id(s)] + hi - lo].copy_(src[lo:hi])
                else:
                    P = sizes[s.group]
                    g0 = starts[id(s)]  # position of the parameter inside the whole-group flat
                    lo, hi = max(g0, rank * P), min(g0 + s.numel, (rank + 1) * P)
                    if lo < hi:
                        dst[lo - rank * P:hi - rank * P].copy_(src[lo - g0:hi - g0])
                        if key == "fp32":
                            mappings[s.group][s.name] = fragment_address(numel=hi - lo, start=lo - rank * P)
            del full
    # torch-optimizer state dict ---------------------------------------------------------------------------------
    from deepspeed_b200.runtime.zero.flat_optimizers import TorchOptimizerAdapter
    if isinstance(zo.flat_opt, TorchOptimizerAdapter):
        inner = {"b200_adapter_state": zo.flat_opt.optimizer.state_dict()}
    else:
        state = {}
        for g in range(n_groups):
            st = {k: outs[k][g] for k in arenas if k != "fp32"}
            st["step"] = torch.tensor(float(zo.group_steps[g] if g < len(zo.group_steps) else 0))
            state[g] = st
        inner = {"state": state, "param_groups": _torch_param_groups(zo)}
    sd = {
        LOSS_SCALER: zo.loss_scaler.state_dict(),
        "dynamic_loss_scale": zo.dynamic_loss_scale,
        "overflow": bool(zo.overflow),
        PARTITION_COUNT: world,
        DS_VERSION: __version__,
        "b200_global_step": zo.global_step,
        "b200_group_steps": list(zo.group_steps),
    }
    if stage3:
        sd[ZERO_STAGE] = 3
        sd[OPTIMIZER_STATE_DICT] = inner
        sd[FP32_FLAT_GROUPS] = outs["fp32"]
    else:
        paddings = []
        for g in range(n_groups):
            left = rank * sizes[g]
            if raw[g] <= left:
                paddings.append(sizes[g])
            elif raw[g] < left + sizes[g]:
                paddings.append(left + sizes[g] - raw[g])
            else:
                paddings.append(0)
        sd[ZERO_STAGE] = 2 if zo.stage == 2 else 1
        sd[CLIP_GRAD] = zo.clip
        sd[BASE_OPTIMIZER_STATE] = inner
        sd[SINGLE_PARTITION_OF_FP32_GROUPS] = [t[:t.numel() - paddings[g]].clone() for g, t in enumerate(outs["fp32"])]
        sd[GROUP_PADDINGS] = paddings
        sd[PARAM_SLICE_MAPPINGS] = mappings
    return sd


def _named_order(zo, param_shapes):
    """Per-group (rt, slot) order: from the checkpoint's ``param_shapes`` (names) when given -- a stock checkpoint orders
    parameters as its optimizer groups did -- else this framework's own order."""
    if not param_shapes:
        return group_param_order(zo)
    by_name = {s.name: (rt, s) for rt in zo.rts for s in rt.u.slots}
    out = []
    for shapes in param_shapes:
        lst = []
        for name, shape in shapes.items():
            if name not in by_name:
                raise KeyError(f"checkpoint parameter '{name}' does not exist in this model")
            rt, s = by_name[name]
            n = 1
            for d in tuple(shape):
                n *= int(d)
            if n != s.numel:
                raise ValueError(f"checkpoint parameter '{name}' has {n} elements, the model's has {s.numel}")
            lst.append((rt, s))
        out.append(lst)
    return out


def _inner_state(sd):
    inner = sd.get(OPTIMIZER_STATE_DICT) if FP32_FLAT_GROUPS in sd else sd.get(BASE_OPTIMIZER_STATE)
    return inner


def _flat_states_by_group(sd, n_groups, stage3):
    """-> (list over groups of {"fp32": flat, "exp_avg": flat, ...}, steps per group)."""
    inner = _inner_state(sd) or {}
    per_group = [dict() for _ in range(n_groups)]
    steps = [None] * n_groups
    if stage3:
        flats = sd[FP32_FLAT_GROUPS]
        # the reference writes one flat per *sub-group* (<= sub_group_size elements, consecutive parameters of one param
        # group); sub-groups of the same param group are concatenated back in order
        pgs = inner.get("param_groups") or []
        sub_to_group = []
        for gi, pg in enumerate(pgs):
            for _ in pg.get("params", []):
                sub_to_group.append(gi)
        if len(sub_to_group) != len(flats):
            sub_to_group = [min(i, n_groups - 1) for i in range(len(flats))] if len(flats) == n_groups else [0] * len(flats)
        st = inner.get("state", {})
        acc = [dict() for _ in range(n_groups)]
        for i, flat in enumerate(flats):
            g = sub_to_group[i]
            acc[g].setdefault("fp32", []).append(flat.float().reshape(-1))
            for k, v in (st.get(i) or {}).items():
                if torch.is_tensor(v) and v.numel() == flat.numel():
                    acc[g].setdefault(k, []).append(v.float().reshape(-1))
                elif k == "step":
                    steps[g] = int(v.item()) if torch.is_tensor(v) else int(v)
        for g in range(n_groups):
            per_group[g] = {k: torch.cat(v) for k, v in acc[g].items()}
            if steps[g] is None and g < len(pgs) and "step" in pgs[g]:
                steps[g] = int(pgs[g]["step"])
    else:
        flats = sd[SINGLE_PARTITION_OF_FP32_GROUPS]
        if isinstance(inner, dict):
            st = inner.get("state", {})
            pgs = inner.get("param_groups") or []
        else:  # elastic checkpoints: a list of lean per-group state dicts
            st = {i: s for i, s in enumerate(inner)}
            pgs = []
        for g in range(min(n_groups, len(flats))):
            per_group[g]["fp32"] = flats[g].float().reshape(-1)
            for k, v in (st.get(g) or {}).items():
                if torch.is_tensor(v) and v.numel() >= flats[g].numel() and v.dim() > 0:
                    per_group[g][k] = v.float().reshape(-1)
                elif k == "step":
                    steps[g] = int(v.item()) if torch.is_tensor(v) else int(v)
            if steps[g] is None and g < len(pgs) and "step" in pgs[g]:
                steps[g] = int(pgs[g]["step"])
    return per_group, steps


@torch.no_grad()
def import_reference_state(zo, sd, load_optimizer_states=True, load_from_fp32_weights=True, param_shapes=None):
    """Scatter a reference-layout shard (ours or stock DeepSpeed's) into the arenas (collective over the DP group)."""
    world, rank = zo.shard_world, zo.shard_rank
    saved_world = sd.get(PARTITION_COUNT, world)
    if isinstance(saved_world, (list, tuple)):
        saved_world = max(saved_world)
    if int(saved_world) != world:
        raise ValueError(f"checkpoint was saved with {saved_world} partitions but this run shards over {world}; convert it "
                         f"with ds_to_universal and load with checkpoint.load_universal")
    stage3 = FP32_FLAT_GROUPS in sd
    order = _named_order(zo, param_shapes)
    n_groups = len(order)
    per_group, steps = _flat_states_by_group(sd, n_groups, stage3)
    arenas = _state_arenas(zo)
    wanted = []
    if load_from_fp32_weights:
        wanted.append("fp32")
    if load_optimizer_states:
        wanted += [k for k in arenas if k != "fp32"]
    missing = [k for k in wanted if any(k not in per_group[g] for g in range(n_groups) if order[g])]
    if missing:
        raise KeyError(f"checkpoint lacks optimizer state tensors {missing}")
    # geometry
    starts, sizes = {}, []
    for g, lst in enumerate(order):
        off = 0
        for _, s in lst:
            starts[id(s)] = off
            off += _partitioned_numel(s.numel, world) if stage3 else s.numel
        if stage3:
            sizes.append(off)
        else:
            align = 2 * world
            sizes.append((-(-off // align) * align) // world)
    group_of = {id(s): g for g, lst in enumerate(order) for _, s in lst}
    for rt in zo.rts:
        members = [s for s in rt.u.slots if id(s) in group_of]
        if not members:
            continue
        u = rt.u
        for key in wanted:
            full = torch.zeros(u.full_numel, dtype=torch.float32)
            for s in members:
                g = group_of[id(s)]
                src = per_group[g][key]
                dst = full[s.offset:s.offset + s.numel]
                if stage3:
                    pn = _partitioned_numel(s.numel, world)
                    lo, hi = rank * pn, min((rank + 1) * pn, s.numel)
                    if lo < hi:
                        dst[lo:hi].copy_(src[starts[id(s)]:starts[id(s)] + hi - lo])
                else:
                    P = sizes[g]
                    g0 = starts[id(s)]
                    lo, hi = max(g0, rank * P), min(g0 + s.numel, (rank + 1) * P)
                    hi = min(hi, rank * P + src.numel())  # lean (unpadded) last partition
                    if lo < hi:
                        dst[lo - g0:hi - g0].copy_(src[lo - rank * P:hi - rank * P])
            if world > 1:
                full = full.to(zo.device)
                dist.all_reduce(full, group=zo.dp_group)
            a = u.arena_offset
            lo_u, hi_u = u.shard_range(rank)
            arena = arenas[key]
            for s in members:  # only the optimizer's own parameters: frozen parameters / padding in the unit keep their values
                lo, hi = max(s.offset, lo_u), min(s.offset + s.numel, hi_u)
                if lo < hi:
                    arena[a + lo - lo_u:a + hi - lo_u].copy_(full[lo:hi].to(arena.device, arena.dtype))
    # scalars -----------------------------------------------------------------------------------------------------------
    ls = sd.get(LOSS_SCALER)
    if isinstance(ls, dict):
        zo.loss_scaler.load_state_dict(ls)
    elif ls is not None:  # a pickled reference LossScaler object
        for attr in ("cur_scale", "cur_iter", "last_overflow_iter", "cur_hysteresis"):
            if hasattr(ls, attr) and hasattr(zo.loss_scaler, attr):
                setattr(zo.loss_scaler, attr, getattr(ls, attr))
    if load_optimizer_states:
        if "b200_group_steps" in sd:
            zo.group_steps = list(sd["b200_group_steps"])
        else:
            zo.group_steps = [int(st if st is not None else 0) for st in steps][:len(zo.group_steps)] + \
                zo.group_steps[len(steps):]
        inner = _inner_state(sd) or {}
        if isinstance(inner, dict) and "b200_adapter_state" in inner:
            zo.flat_opt.optimizer.load_state_dict(inner["b200_adapter_state"])
        for g, saved in zip(zo.param_groups, (inner.get("param_groups") if isinstance(inner, dict) else None) or []):
            for k, v in saved.items():
                if k not in ("params", "step") and k in g:
                    g[k] = v
    zo.global_step = int(sd.get("b200_global_step", max([s for s in steps if s is not None] or [0])))
    if load_from_fp32_weights:
        if zo.master is None:  # fp32 training: "fp32" arena was a temporary concatenation of the lp shards
            flat = arenas["fp32"]
            for rt in zo.rts:
                a0 = rt.u.arena_offset
                zo._lp_shard(rt.u).copy_(flat[a0:a0 + rt.u.shard_numel])
        zo._refresh_lp_from_master()


@torch.no_grad()
def import_reference_state_elastic(zo, shards, load_optimizer_states=True, load_from_fp32_weights=True, param_shapes=None):
    """Load a reference-layout checkpoint written with a DIFFERENT data-parallel degree (the reference's
    ``elastic_checkpoint`` behaviour, ``stage_1_and_2.py:2290 _restore_elastic_base_optimizer_state`` -- extended here to
    stage 3): ``shards`` are the optimizer shards of ALL saved DP ranks (in rank order); every parameter's fp32 value and
    optimizer moments are reassembled from them and scattered into this run's arenas.  Host memory: the full fp32 model +
    moments once per rank while loading, like the reference's elastic path."""
    from deepspeed_b200.checkpoint.ds_to_universal import reference_param_states
    order = _named_order(zo, param_shapes)
    shapes = [OrderedDict((s.name, tuple(s.shape)) for _, s in lst) for lst in order]
    params, steps, _ = reference_param_states(shards, shapes)
    arenas = _state_arenas(zo)
    wanted = (["fp32"] if load_from_fp32_weights else []) + ([k for k in arenas if k != "fp32"] if load_optimizer_states else [])
    for lst in order:
        for rt, s in lst:
            st = params[s.name]
            for key in wanted:
                if key not in st:
                    raise KeyError(f"checkpoint lacks optimizer state '{key}' for parameter '{s.name}'")
                zo._scatter_into_arena(arenas[key], rt, s, st[key])
    sd = shards[0]
    ls = sd.get(LOSS_SCALER)
    if isinstance(ls, dict):
        zo.loss_scaler.load_state_dict(ls)
    if load_optimizer_states:
        if "b200_group_steps" in sd:
            zo.group_steps = list(sd["b200_group_steps"])
        else:
            zo.group_steps = [int(st or 0) for st in steps][:len(zo.group_steps)] + zo.group_steps[len(steps):]
        for g, saved in zip(zo.param_groups, (_inner_state(sd) or {}).get("param_groups", []) if isinstance(_inner_state(sd), dict) else []):
            for k, v in saved.items():
                if k not in ("params", ) and k in g:
                    g[k] = v
    if load_from_fp32_weights:
        zo._refresh_lp_from_master()
