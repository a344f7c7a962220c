#!/usr/bin/env python
"""Convert a ZeRO checkpoint into the *universal* per-parameter format (any DP/TP/PP world can load it).

Reference: ``checkpoint/ds_to_universal.py`` (extract :112/:152, merge :232) and the on-disk layout
``<out>/zero/<param_name>/{fp32.pt, exp_avg.pt, exp_avg_sq.pt, step.pt}`` + ``<out>/mp_rank_00_model_states.pt``.
Because this framework's shards are plain contiguous slices of unit-flat buffers, extraction and merge are a
single pass (concatenate the rank slices of a unit, cut parameters out) and need no temp directory.
"""
import argparse
import os
import shutil

import torch

from deepspeed_b200.utils import zero_to_fp32 as Z
from .constants import UNIVERSAL_CHECKPOINT_INFO, UNIVERSAL_CHECKPOINT_VERSION_KEY, UNIVERSAL_CHECKPOINT_VERSION_VALUE


def convert_to_universal(input_folder, output_folder, tag=None, keep_temp_folder=False, num_extract_workers=1,
                         num_merge_workers=1, inject_missing_state=False, strict=True):
    ds_dir = Z._resolve_tag(input_folder, tag) if (tag is not None or os.path.isfile(os.path.join(input_folder, "latest"))) \
        else input_folder
    ms = Z._model_state(ds_dir)
    by_mp = Z.get_optim_shards(ds_dir)
    if len(by_mp) > 1:
        raise NotImplementedError("TP-sharded checkpoints: convert each mp rank separately (tp merge uses "
                                  "universal_checkpoint_info patterns)")
    shards = [Z._load(f)["optimizer_state_dict"] for f in by_mp[0]]
    zero_dir = os.path.join(output_folder, "zero")
    os.makedirs(zero_dir, exist_ok=True)
    if "fp32_flat" not in shards[0]:
        # the reference's on-disk layout (stock DeepSpeed, or this framework's default checkpoint.b200_shard_layout)
        return _convert_reference_layout(ms, shards, output_folder, zero_dir)
    layout = ms["ds_b200_layout"]
    state_names = list(shards[0].get("flat_state", {}).keys())
    step = shards[0].get("group_steps", [0])
    flats = {"fp32": list(Z._unit_flats(layout, shards))}
    for sn in state_names:
        flats[sn] = list(Z._unit_flats(layout, shards, key="flat_state", sub=sn))
    n = 0
    for ui, (u, _) in enumerate(flats["fp32"]):
        for (name, off, numel, shape, group) in u["slots"]:
            d = os.path.join(zero_dir, name)
            os.makedirs(d, exist_ok=True)
            for key, lst in flats.items():
                t = lst[ui][1][off:off + numel].view(*shape).clone()
                torch.save({"param": t}, os.path.join(d, f"{key}.pt"))
            g = group if isinstance(group, int) and 0 <= group < len(step) else 0
            torch.save(torch.tensor(float(step[g] if step else 0)), os.path.join(d, "step.pt"))
            n += 1
    ms = dict(ms)
    ms[UNIVERSAL_CHECKPOINT_INFO] = {UNIVERSAL_CHECKPOINT_VERSION_KEY: UNIVERSAL_CHECKPOINT_VERSION_VALUE}
    ms["optimizer_meta"] = {k: shards[0].get(k) for k in ("param_groups", "group_steps", "global_step", "loss_scaler",
                                                         "zero_stage")}
    torch.save(ms, os.path.join(output_folder, "mp_rank_00_model_states.pt"))
    parent = os.path.dirname(os.path.normpath(output_folder))
    with open(os.path.join(parent, "latest_universal"), "w") as f:
        f.write(os.path.basename(os.path.normpath(output_folder)))
    print(f"universal checkpoint: {n} parameters x {1 + len(state_names)} tensors -> {output_folder}")
    return output_folder


def reference_param_states(shards, param_shapes):
    """Per-parameter tensors out of reference-layout optimizer shards (all DP ranks of one mp rank).

    -> ``(OrderedDict name -> {"fp32": t, "exp_avg": t, ...} (param-shaped), steps per group, inner param_groups)``.
    Stage 3: each parameter contributes ``ceil(numel / world)`` elements to every rank's flat (sub-groups are consecutive
    runs of parameters, so concatenating a rank's sub-group flats restores the walk).  Stage 1/2: the rank partitions of a
    group concatenate to the group's flat, parameters sit back to back in it."""
    from collections import OrderedDict
    world = len(shards)
    stage3 = "fp32_flat_groups" in shards[0]
    inner0 = shards[0].get("optimizer_state_dict" if stage3 else "base_optimizer_state") or {}
    pgs = inner0.get("param_groups", []) if isinstance(inner0, dict) else []
    out = OrderedDict()
    steps = []

    def states_of(sd, idx):
        inner = sd.get("optimizer_state_dict" if stage3 else "base_optimizer_state") or {}
        st = inner.get("state", {}) if isinstance(inner, dict) else {i: x for i, x in enumerate(inner)}
        return st.get(idx) or {}

    if stage3:
        per_rank = []
        for sd in shards:
            flats = sd["fp32_flat_groups"]
            acc = {"fp32": [f.float().reshape(-1) for f in flats]}
            for i, f in enumerate(flats):
                for k, v in states_of(sd, i).items():
                    if torch.is_tensor(v) and v.numel() == f.numel():
                        acc.setdefault(k, []).append(v.float().reshape(-1))
            per_rank.append({k: torch.cat(v) for k, v in acc.items()})
        st0 = states_of(shards[0], 0)
        step0 = st0.get("step", pgs[0].get("step", 0) if pgs else 0)
        steps = [float(step0) for _ in (param_shapes or [None])]
        off = 0
        for shapes in param_shapes:
            for name, shape in shapes.items():
                n = Z._numel(shape)
                per = -(-n // world)
                out[name] = {k: torch.cat([pr[k][off:off + per] for pr in per_rank])[:n].view(*shape).clone()
                             for k in per_rank[0]}
                off += per
    else:
        for g, shapes in enumerate(param_shapes):
            keys = {"fp32": [sd["single_partition_of_fp32_groups"][g].float().reshape(-1) for sd in shards]}
            P = None
            for r, sd in enumerate(shards):
                for k, v in states_of(sd, g).items():
                    if torch.is_tensor(v) and v.dim() > 0 and v.numel() > 1:
                        keys.setdefault(k, []).append(v.float().reshape(-1))
            st0 = states_of(shards[0], g)
            steps.append(float(st0.get("step", pgs[g].get("step", 0) if g < len(pgs) else 0)))
            full = {k: torch.cat(v) for k, v in keys.items()}
            off = 0
            for name, shape in shapes.items():
                n = Z._numel(shape)
                out[name] = {k: t[off:off + n].view(*shape).clone() for k, t in full.items()}
                off += n
    return out, steps, pgs


def _convert_reference_layout(ms, shards, output_folder, zero_dir):
    params, steps, pgs = reference_param_states(shards, ms["param_shapes"])
    group_of = {name: g for g, shapes in enumerate(ms["param_shapes"]) for name in shapes}
    for name, states in params.items():
        d = os.path.join(zero_dir, name)
        os.makedirs(d, exist_ok=True)
        for key, t in states.items():
            torch.save({"param": t}, os.path.join(d, f"{key}.pt"))
        g = group_of.get(name, 0)
        torch.save(torch.tensor(float(steps[g] if g < len(steps) else 0)), os.path.join(d, "step.pt"))
    ms = dict(ms)
    ms[UNIVERSAL_CHECKPOINT_INFO] = {UNIVERSAL_CHECKPOINT_VERSION_KEY: UNIVERSAL_CHECKPOINT_VERSION_VALUE}
    s0 = shards[0]
    ls = s0.get("loss_scaler")
    ms["optimizer_meta"] = {
        "param_groups": [{k: v for k, v in g.items() if k not in ("params", "step")} for g in pgs],
        "group_steps": s0.get("b200_group_steps", [int(x) for x in steps]),
        "global_step": s0.get("b200_global_step", int(max(steps) if steps else 0)),
        "loss_scaler": ls if isinstance(ls, dict) else None,
        "zero_stage": s0.get("zero_stage"),
    }
    torch.save(ms, os.path.join(output_folder, "mp_rank_00_model_states.pt"))
    parent = os.path.dirname(os.path.normpath(output_folder))
    with open(os.path.join(parent, "latest_universal"), "w") as f:
        f.write(os.path.basename(os.path.normpath(output_folder)))
    n_states = len(next(iter(params.values()))) if params else 0
    print(f"universal checkpoint: {len(params)} parameters x {n_states} tensors -> {output_folder}")
    return output_folder


# ---- staged pipeline (extract per-rank fragments → merge) ----------------------------------------------------------------
# The one-pass converter above holds a whole unit in memory.  For checkpoints where that is not possible -- and for
# checkpoints in the upstream DeepSpeed layout -- the same result is produced in two restartable stages that mirror the
# reference tool (``ds_to_universal.py:112-346``): every DP rank's optimizer file is cut into per-parameter fragment files
# ``<tmp>/<param>/<tp>/<state>.<dp>``, then the fragments of one parameter are concatenated (over DP) and merged over TP by
# the rules in ``universal_checkpoint_info``.
_STATES = ("fp32", "exp_avg", "exp_avg_sq")


def atoi(text):
    return Z.atoi(text)


def natural_keys(text):
    return Z.natural_keys(text)


def dp_index_to_str(dp_index):
    return f"{dp_index:0>2d}"


def dump_param_fragment(dir, tp_index, dp_index, state_name, state_flat_tensor, param_name, offset, numel):
    """Write ``state_flat_tensor[offset:offset+numel]`` (or the scalar as is) to ``<dir>/<param>/<tp>/<state>.<dp>``."""
    base = os.path.join(dir, param_name, str(tp_index))
    os.makedirs(base, exist_ok=True)
    val = state_flat_tensor
    if state_name != "step" and torch.is_tensor(val):
        val = val.narrow(0, offset, numel).clone()
    torch.save(val, os.path.join(base, f"{state_name}.{dp_index_to_str(dp_index)}"))


def extract_zero_shards(dir, ds_checkpoint, indices_3D):
    """Stage-1/2 extraction of one (pp, tp, dp) rank.  ``ds_checkpoint`` is a ``DeepSpeedCheckpoint``; its optimizer state
    must carry ``param_slice_mappings`` (per group: name → fragment with ``start`` / ``numel``), the flat fp32 partition
    per group, and Adam moments per group."""
    pp, tp, dp = indices_3D
    osd = ds_checkpoint.get_zero_checkpoint_state(pp_index=pp, tp_index=tp, dp_index=dp)["optimizer_state_dict"]
    mappings = osd["param_slice_mappings"]
    fp32_groups = osd["single_partition_of_fp32_groups"]
    inner = osd.get("base_optimizer_state") or osd.get("optimizer_state_dict", {}).get("state", {})
    pipeline_replicated = ds_checkpoint.get_checkpoint_info().get("pipeline_replicated_parameter_patterns", []) \
        if hasattr(ds_checkpoint, "get_checkpoint_info") else []
    import re
    for g, frags in enumerate(mappings):
        st = inner[g] if not isinstance(inner, dict) else inner.get(g, {})
        flat = {"fp32": fp32_groups[g]}
        for k in ("exp_avg", "exp_avg_sq", "step"):
            if k in st:
                flat[k] = st[k]
        for name, frag in frags.items():
            if pp > 0 and any(re.match(pat, name) for pat in pipeline_replicated):
                continue  # tied copies live on the first stage only
            start = frag["start"] if isinstance(frag, dict) else frag.start
            numel = frag["numel"] if isinstance(frag, dict) else frag.numel
            for key, t in flat.items():
                dump_param_fragment(dir, tp, dp, key, t, name, start, numel)


def _partition_info(numel, world):
    per = -(-numel // world)
    return per, per * world - numel


def extract_zero_shards_stage3(optim_files, param_shapes, dp_degree, temp_dir, dp_index):
    """Stage-3 extraction of DP rank ``dp_index`` from the upstream layout: parameters sit back to back in the rank's flat
    group, each contributing ``ceil(numel/dp)`` elements (the tail rank's share is shorter: padding is dropped here)."""
    osd = torch.load(optim_files[dp_index], map_location="cpu", weights_only=False)["optimizer_state_dict"]
    inner = osd["optimizer_state_dict"]["state"][0]
    flat = {"exp_avg": inner["exp_avg"], "exp_avg_sq": inner["exp_avg_sq"], "fp32": osd["fp32_flat_groups"][0]}
    off = 0
    for name, shape in param_shapes.items():
        n = Z._numel(shape)
        per, _ = _partition_info(n, dp_degree)
        real = max(0, min(per, n - dp_index * per))
        for key, t in flat.items():
            dump_param_fragment(temp_dir, 0, dp_index, key, t, name, off, real)
        off += per


def _merge_zero_shards(param_base_path, state, tp_degree, slice_shape=None):
    """Concatenate the DP fragments of ``state`` for every TP index; returns one tensor per TP slice."""
    import glob
    out = []
    for tp in range(tp_degree):
        paths = sorted(glob.glob(os.path.join(param_base_path, str(tp), f"{state}.*")), key=natural_keys)
        if not paths:
            continue
        pieces = [torch.load(p, map_location="cpu", weights_only=False) for p in paths]
        if state == "step":
            assert all(float(v) == float(pieces[0]) for v in pieces), "all fragments must have the same step value"
            out.append(pieces[0])
            continue
        t = torch.cat([x.reshape(-1) for x in pieces])
        out.append(t.view(slice_shape) if slice_shape is not None else t)
    return out


def merge_zero3_slices(dp_degree, dir, slice_dir, name):
    for state in _STATES:
        merged = _merge_zero_shards(os.path.join(slice_dir, name), state, 1)
        os.makedirs(os.path.join(dir, name), exist_ok=True)
        torch.save({"param": merged[0]}, os.path.join(dir, name, f"{state}.pt"))


def merge_tp_slices(ds_checkpoint, dir, slice_dir, tp_degree, name_and_shape):
    """Merge the TP slices of one parameter by the checkpoint's ``universal_checkpoint_info`` rules (replicated / averaged /
    row-parallel / fused sub-parameters / vocabulary padding).  Returns the patterns that matched nothing (for ``--strict``)."""
    import re
    from . import constants as K
    from .universal_checkpoint import SubparamShape
    name, shape = name_and_shape
    info = ds_checkpoint.get_checkpoint_info(K.UNIVERSAL_CHECKPOINT_INFO) if hasattr(ds_checkpoint, "get_checkpoint_info") \
        else dict(ds_checkpoint)
    rules = {k: list(info.get(k, [])) for k in (K.TP_REPLICATED_PARAMETER_PATTERNS, K.PARAMETER_TO_AVERAGE_PATTERNS,
                                                K.PARAMETER_WITH_ROW_PARALLELISM_PATTERNS, K.VOCABULARY_PARAMETER_PATTERNS,
                                                K.PARAMETER_WITH_2_SUB_PARAMS_CAT_DIM_0)}
    subs = [SubparamShape(**d) if isinstance(d, dict) else d for d in info.get(K.PARAMETER_WITH_SUB_PARAMS, [])]
    unmatched = {p for ps in rules.values() for p in ps} | {p for sp in subs for p in sp.patterns}

    def hit(kind):
        m = [p for p in rules[kind] if re.match(p, name)]
        assert len(m) <= 1, f"Got more than one matching patterns={m} for {name}"
        if m:
            unmatched.discard(m[0])
        return bool(m)

    sub = None
    for sp in subs:
        for p in sp.patterns:
            if re.match(p, name):
                unmatched.discard(p)
                sub = sp
    src, dst = os.path.join(slice_dir, name), os.path.join(dir, name)
    os.makedirs(dst, exist_ok=True)
    step = _merge_zero_shards(src, "step", tp_degree, shape)
    if step:
        torch.save(step[0], os.path.join(dst, "step.pt"))
    for state in _STATES:
        sl = _merge_zero_shards(src, state, tp_degree, shape)
        if not sl:
            continue
        rec = {}
        if hit(K.TP_REPLICATED_PARAMETER_PATTERNS):
            assert all(sl[0].equal(o) for o in sl[1:]), f"{name}: replicated parameter differs across TP ranks"
            full = sl[0]
        elif hit(K.PARAMETER_TO_AVERAGE_PATTERNS):
            full = sum(sl) / len(sl)
        elif hit(K.PARAMETER_WITH_2_SUB_PARAMS_CAT_DIM_0):
            halves = [s.chunk(2, dim=0) for s in sl]
            full = torch.cat([h[0] for h in halves] + [h[1] for h in halves], dim=0)
            rec[K.CAT_DIM], rec[K.PARAM_N_SUB_PARAMS] = 0, 2
        elif sub is not None:
            pd = sub.partition_dim
            sizes = sub.shape[pd] if isinstance(sub.shape[pd], tuple) else (sub.shape[pd], )
            local = [sum(d) if isinstance(d, tuple) else d for d in sub.shape]
            local[pd] //= tp_degree
            views = [s.view(local) for s in sl]
            cols, at = [], 0
            for sz in sizes:
                w = sz // tp_degree
                cols.append(torch.cat([v.narrow(pd, at, w) for v in views], dim=pd))
                at += w
            full = torch.cat(cols, dim=pd)
            rec[K.SUB_PARAM_SHAPE] = sub
        else:
            rec[K.CAT_DIM] = 1 if hit(K.PARAMETER_WITH_ROW_PARALLELISM_PATTERNS) else 0
            full = torch.cat(sl, dim=rec[K.CAT_DIM])
        if hit(K.VOCABULARY_PARAMETER_PATTERNS):
            full = full[:info[K.ORIGINAL_VOCAB_SIZE], :]
            rec[K.VOCAB_TENSOR] = True
        rec["param"] = full
        torch.save(rec, os.path.join(dst, f"{state}.pt"))
    return unmatched


def convert_upstream_stage3_to_universal(input_folder, output_folder, keep_temp_folder=False):
    """Upstream-layout ZeRO-3 checkpoint → universal (extract + merge per parameter)."""
    optim_files = Z.get_optim_files(input_folder)
    dp = len(optim_files)
    states = Z.parse_model_states(Z.get_model_state_files(input_folder)[:1])
    shapes = {k: v for d in states[0].param_shapes for k, v in d.items()}
    tmp, zero_dir = os.path.join(output_folder, "tmp"), os.path.join(output_folder, "zero")
    os.makedirs(tmp, exist_ok=True)
    for r in range(dp):
        extract_zero_shards_stage3(optim_files, shapes, dp, tmp, r)
    for name, shape in shapes.items():
        merge_zero3_slices(dp, zero_dir, tmp, name)
        for st in _STATES:  # reshape the flat merge to the parameter's shape
            f = os.path.join(zero_dir, name, f"{st}.pt")
            rec = torch.load(f, weights_only=False)
            rec["param"] = rec["param"][:Z._numel(shape)].view(shape)
            torch.save(rec, f)
    ms = Z._load(Z.get_model_state_files(input_folder)[0])
    ms[UNIVERSAL_CHECKPOINT_INFO] = {UNIVERSAL_CHECKPOINT_VERSION_KEY: UNIVERSAL_CHECKPOINT_VERSION_VALUE}
    torch.save(ms, os.path.join(output_folder, "mp_rank_00_model_states.pt"))
    if not keep_temp_folder:
        shutil.rmtree(tmp, ignore_errors=True)
    return output_folder


def parse_arguments(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--input_folder", type=str, required=True, help="<save_dir>/<tag> of the ZeRO checkpoint")
    ap.add_argument("--output_folder", type=str, required=True)
    ap.add_argument("--num_extract_workers", default=4, type=int)
    ap.add_argument("--num_merge_workers", default=2, type=int)
    ap.add_argument("--keep_temp_folder", action="store_true")
    ap.add_argument("--no_strict", dest="strict", action="store_false")
    ap.add_argument("--inject_missing_state", action="store_true")
    return ap.parse_args(argv)


def main(args=None):
    a = args if args is not None else parse_arguments()
    ms = Z._model_state(a.input_folder)
    if "ds_b200_layout" not in ms and "param_shapes" in ms:
        osd = Z._load(Z.get_optim_files(a.input_folder)[0])["optimizer_state_dict"]
        if osd.get("zero_stage") == 3 and len(osd.get("fp32_flat_groups", [])) == 1:
            return convert_upstream_stage3_to_universal(a.input_folder, a.output_folder, a.keep_temp_folder)
    return convert_to_universal(a.input_folder, a.output_folder, strict=a.strict,
                                inject_missing_state=a.inject_missing_state)


if __name__ == "__main__":
    main()
