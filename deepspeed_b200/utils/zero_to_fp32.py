#!/usr/bin/env python
"""Consolidate a ZeRO (stage 1/2/3) checkpoint written by deepspeed_b200 into a plain fp32 ``state_dict``.

This script is copied next to every checkpoint (like the reference's ``utils/zero_to_fp32.py``) and depends on
``torch`` only.  Usage::

    python zero_to_fp32.py <checkpoint_dir> <output_dir> [--tag TAG] [--max_shard_size 5GB] [--safe_serialization]

API parity: ``get_fp32_state_dict_from_zero_checkpoint``, ``convert_zero_checkpoint_to_fp32_state_dict``,
``load_state_dict_from_zero_checkpoint``.

Checkpoints written by upstream DeepSpeed (``param_shapes`` + ``single_partition_of_fp32_groups`` / ``fp32_flat_groups``)
are recognised too and consolidated by the same entry points, so existing runs can be converted with this tool.

Layout read (see ``runtime/checkpointing.py``): every DP rank's ``*_optim_states.pt`` holds ``fp32_flat`` — its
slice of each *unit* (unit u occupies ``[arena_offset, arena_offset+shard_numel)``); the unit's full flat
buffer is the rank-order concatenation of those slices and ``ds_b200_layout`` in the model-states file lists
each parameter's ``(name, offset, numel, shape)`` inside it.
"""
import argparse
import glob
import json
import os
import re
from collections import OrderedDict

import torch


def atoi(text):
    return int(text) if text.isdigit() else text


def natural_keys(text):
    """Sort key under which ``rank_10`` follows ``rank_9``."""
    return [atoi(t) for t in re.split(r"(\d+)", text)]


_natural = natural_keys


def _resolve_tag(checkpoint_dir, tag):
    if tag is None:
        latest = os.path.join(checkpoint_dir, "latest")
        if not os.path.isfile(latest):
            raise ValueError(f"Unable to find 'latest' file at {latest}")
        with open(latest) as f:
            tag = f.read().strip()
    d = os.path.join(checkpoint_dir, tag)
    if not os.path.isdir(d):
        raise FileNotFoundError(f"Directory '{d}' doesn't exist")
    return d


def _load(path):
    return torch.load(path, map_location="cpu", weights_only=False)


def _files(ds_dir, pattern):
    fs = sorted(glob.glob(os.path.join(ds_dir, pattern)), key=_natural)
    if not fs:
        raise FileNotFoundError(f"can't find {pattern} files in directory '{ds_dir}'")
    return fs


def _model_state(ds_dir):
    cands = sorted(glob.glob(os.path.join(ds_dir, "*_model_states.pt")), key=_natural)
    cands = [c for c in cands if "expert_" not in os.path.basename(c)]
    if not cands:
        raise FileNotFoundError(f"no *_model_states.pt under {ds_dir}")
    return _load(cands[0])


def get_optim_shards(ds_dir):
    files = _files(ds_dir, "*_optim_states.pt")
    # one file per DP rank (mp_rank_00 only: TP-sharded checkpoints are consolidated per mp rank)
    by_mp = {}
    for f in files:
        m = re.search(r"zero_pp_rank_(\d+)_mp_rank_(\d+)_optim_states", os.path.basename(f))
        by_mp.setdefault(int(m.group(2)), []).append((int(m.group(1)), f))
    return {mp: [f for _, f in sorted(v)] for mp, v in by_mp.items()}


def _unit_flats(layout, shards, key="fp32_flat", sub=None):
    """Yield (unit_dict, full_flat_fp32) by concatenating every rank's slice."""
    for u in layout["units"]:
        a, n = u["arena_offset"], u["shard_numel"]
        parts = []
        for sd in shards:
            src = sd[key] if sub is None else sd[key][sub]
            parts.append(src[a:a + n].float())
        yield u, torch.cat(parts)


# ---- upstream-DeepSpeed checkpoint layout -------------------------------------------------------------------------------
class zero_model_state:
    """What a ``*_model_states.pt`` of the upstream layout contributes to consolidation."""

    def __init__(self, buffers, param_shapes, shared_params, ds_version, frozen_param_shapes, frozen_param_fragments):
        self.buffers, self.param_shapes, self.shared_params, self.ds_version = buffers, param_shapes, shared_params, ds_version
        self.frozen_param_shapes, self.frozen_param_fragments = frozen_param_shapes, frozen_param_fragments


def get_checkpoint_files(checkpoint_dir, glob_pattern):
    return _files(checkpoint_dir, glob_pattern)


def get_optim_files(checkpoint_dir):
    return get_checkpoint_files(checkpoint_dir, "*_optim_states.pt")


def get_model_state_files(checkpoint_dir):
    return get_checkpoint_files(checkpoint_dir, "*_model_states.pt")


def get_model_state_file(checkpoint_dir, zero_stage):
    if not os.path.isdir(checkpoint_dir):
        raise FileNotFoundError(f"Directory '{checkpoint_dir}' doesn't exist")
    name = "mp_rank_00_model_states.pt" if zero_stage <= 2 else "zero_pp_rank_0_mp_rank_00_model_states.pt"
    f = os.path.join(checkpoint_dir, name)
    if not os.path.exists(f):
        raise FileNotFoundError(f"can't find model states file at '{f}'")
    return f


def parse_model_states(files):
    out = []
    for f in files:
        sd = _load(f)
        if "buffer_names" not in sd:
            raise ValueError(f"{f} is not a model state checkpoint")
        names = set(sd["buffer_names"])
        out.append(zero_model_state(buffers={k: v.float() for k, v in sd["module"].items() if k in names},
                                    param_shapes=sd["param_shapes"],
                                    shared_params=[[k, v] for k, v in (sd.get("shared_params") or {}).items()],
                                    ds_version=sd.get("ds_version"), frozen_param_shapes=sd.get("frozen_param_shapes"),
                                    frozen_param_fragments=sd.get("frozen_param_fragments")))
    return out


def parse_optim_states(files, ds_checkpoint_dir):
    """-> ``(zero_stage, world_size, per-rank list of fp32 flat groups)``."""
    osds = []
    for f in files:
        osd = torch.load(f, map_location="cpu", mmap=True, weights_only=False)["optimizer_state_dict"]
        osd.pop("optimizer_state_dict", None)  # moments are not needed, only the fp32 master weights
        osds.append(osd)
    if "zero_stage" not in osds[0]:
        raise ValueError(f"{files[0]} is not a zero checkpoint")
    stage, world = osds[0]["zero_stage"], osds[0]["partition_count"]
    world = max(world) if isinstance(world, (list, tuple)) else world
    if world != len(files):
        raise ValueError(f"Expected {world} of '*_optim_states.pt' under '{ds_checkpoint_dir}' but found {len(files)} files. "
                         "Possibly due to an overwrite of an old checkpoint, or a checkpoint didn't get saved by one or "
                         "more processes.")
    key = {1: "single_partition_of_fp32_groups", 2: "single_partition_of_fp32_groups", 3: "fp32_flat_groups"}.get(stage)
    if key is None:
        raise ValueError(f"unknown zero stage {stage}")
    return stage, world, [o[key] for o in osds]


def _numel(shape):
    n = 1
    for d in shape:
        n *= int(d)
    return n


def zero3_partitioned_param_info(unpartitioned_numel, world_size):
    """(elements each rank holds, padding on the last rank) of a stage-3 parameter."""
    per = -(-unpartitioned_numel // world_size)
    return per, per * world_size - unpartitioned_numel


class GatheredTensor:
    """Deferred stage-3 parameter: remembers where its slice sits in every rank's (virtually concatenated) flat groups and
    assembles the full tensor only on ``contiguous()`` so consolidation can stream shard by shard."""

    def __init__(self, flat_groups, flat_groups_offset, offset, partitioned_numel, shape):
        self.flat_groups, self.flat_groups_offset = flat_groups, flat_groups_offset
        self.offset, self.partitioned_numel, self.shape = offset, partitioned_numel, torch.Size(shape)
        self.dtype = flat_groups[0][0].dtype

    def _rank_slice(self, groups):
        lo, hi = self.offset, self.offset + self.partitioned_numel
        bounds = self.flat_groups_offset
        pieces = []
        for g, t in enumerate(groups):
            a, b = max(lo, bounds[g]), min(hi, bounds[g + 1])
            if a < b:
                pieces.append(t[a - bounds[g]:b - bounds[g]])
        return pieces

    def contiguous(self):
        chunks = [c for groups in self.flat_groups for c in self._rank_slice(groups)]
        return torch.cat(chunks)[:self.shape.numel()].view(self.shape).contiguous()

    def numel(self):
        return self.shape.numel()


def to_torch_tensor(state_dict, return_empty_tensor=False):
    """Materialise the ``GatheredTensor`` entries (shared entries materialised once)."""
    done, out = {}, {}
    for name, t in state_dict.items():
        if not isinstance(t, GatheredTensor):
            out[name] = t
        elif id(t) in done:
            out[name] = out[done[id(t)]]
        else:
            done[id(t)] = name
            out[name] = torch.empty(t.shape, dtype=t.dtype) if return_empty_tensor else t.contiguous()
    return out


def _consolidate_upstream(ds_dir, exclude_frozen_parameters, lazy_mode):
    stage, world, flats = parse_optim_states(get_optim_files(ds_dir), ds_dir)
    print(f"Detected checkpoint of type zero stage {stage}, world_size: {world}")
    states = parse_model_states(get_model_state_files(ds_dir))
    print(f"Parsing checkpoint created by deepspeed=={states[0].ds_version}")
    out = OrderedDict(states[0].buffers)
    frozen = states[0].frozen_param_shapes or {}
    if frozen and not exclude_frozen_parameters:
        for name, shape in frozen.items():
            if stage <= 2:
                out[name] = states[0].frozen_param_fragments[name]
            else:
                frags = [st.frozen_param_fragments[name] for st in states]
                out[name] = torch.cat(frags)[:_numel(shape)].view(shape)
    if stage <= 2:
        align = 2 * world
        for g, shapes in enumerate(states[0].param_shapes):
            full = torch.cat([rank_groups[g] for rank_groups in flats])
            off = 0
            for name, shape in shapes.items():
                n = _numel(shape)
                out[name] = full[off:off + n].view(shape)
                off += n
            up = lambda x: -(-x // align) * align
            if up(off) != up(full.numel()):
                raise ValueError(f"consumed {off} numels out of {full.numel()} - something is wrong")
    else:
        shapes = {k: v for d in states[0].param_shapes for k, v in d.items()}
        bounds = [0]
        for t in flats[0]:
            bounds.append(bounds[-1] + t.numel())
        off = 0
        for name, shape in shapes.items():
            per, _ = zero3_partitioned_param_info(_numel(shape), world)
            out[name] = GatheredTensor(flats, bounds, off, per, shape)
            off += per
        if off * world != bounds[-1] * world:
            raise ValueError(f"consumed {off * world} numels out of {bounds[-1] * world} - something is wrong")
    for alias, canon in states[0].shared_params:
        if canon in out:
            out[alias] = out[canon]
    return out if lazy_mode else to_torch_tensor(out)


def get_fp32_state_dict_from_zero_checkpoint(checkpoint_dir, tag=None, exclude_frozen_parameters=False, lazy_mode=False):
    ds_dir = _resolve_tag(checkpoint_dir, tag)
    ms = _model_state(ds_dir)
    layout = ms.get("ds_b200_layout")
    first = _load(get_optim_shards(ds_dir)[0][0])["optimizer_state_dict"]
    if layout is None or "fp32_flat" not in first:
        # upstream on-disk layout: written by stock DeepSpeed, or by this framework with checkpoint.b200_shard_layout =
        # "reference" (the default)
        if "param_shapes" in ms:
            return _consolidate_upstream(ds_dir, exclude_frozen_parameters, lazy_mode)
        raise ValueError("checkpoint has neither ds_b200_layout nor param_shapes: not a ZeRO checkpoint")
    del first
    shards = [_load(f)["optimizer_state_dict"] for f in get_optim_shards(ds_dir)[0]]
    world = shards[0]["partition_count"]
    if len(shards) != world:
        raise ValueError(f"Expected {world} of '*_optim_states.pt' under '{ds_dir}' but found {len(shards)} files")
    print(f"Detected checkpoint of type zero stage {shards[0]['zero_stage']}, world_size: {world}")
    frozen = set((ms.get("frozen_param_shapes") or {}).keys())
    out = OrderedDict()
    buffers = set(ms.get("buffer_names") or [])
    for k, v in ms["module"].items():
        if k in buffers:
            out[k] = v.float() if v.is_floating_point() else v
    for u, flat in _unit_flats(layout, shards):
        for (name, off, numel, shape, group) in u["slots"]:
            if exclude_frozen_parameters and name in frozen:
                continue
            out[name] = flat[off:off + numel].view(*shape).clone()
    # parameters that are not managed by ZeRO (stage<=2 keeps none outside; ZeRO-3 external) fall back to module
    for k, v in ms["module"].items():
        if k not in out and v.numel() > 0:
            out[k] = v.float() if v.is_floating_point() else v
    for alias, canon in (ms.get("shared_params") or {}).items():
        if canon in out:
            out[alias] = out[canon]
    return out


def _parse_size(s):
    if isinstance(s, int):
        return s
    m = re.match(r"^(\d+(?:\.\d+)?)\s*([KMGT]?B)$", s.strip().upper())
    if not m:
        raise ValueError(f"bad size {s}")
    return int(float(m.group(1)) * {"B": 1, "KB": 10**3, "MB": 10**6, "GB": 10**9, "TB": 10**12}[m.group(2)])


def convert_zero_checkpoint_to_fp32_state_dict(checkpoint_dir, output_dir, max_shard_size="5GB", safe_serialization=False,
                                               tag=None, exclude_frozen_parameters=False):
    sd = get_fp32_state_dict_from_zero_checkpoint(checkpoint_dir, tag, exclude_frozen_parameters)
    os.makedirs(output_dir, exist_ok=True)
    limit = _parse_size(max_shard_size)
    shards, cur, cur_bytes = [], OrderedDict(), 0
    for k, v in sd.items():
        b = v.numel() * v.element_size()
        if cur and cur_bytes + b > limit:
            shards.append(cur)
            cur, cur_bytes = OrderedDict(), 0
        cur[k] = v
        cur_bytes += b
    shards.append(cur)
    ext = "safetensors" if safe_serialization else "bin"
    base = "model" if safe_serialization else "pytorch_model"
    index = {"metadata": {"total_size": sum(v.numel() * v.element_size() for v in sd.values())}, "weight_map": {}}
    for i, sh in enumerate(shards):
        fn = f"{base}.{ext}" if len(shards) == 1 else f"{base}-{i + 1:05d}-of-{len(shards):05d}.{ext}"
        path = os.path.join(output_dir, fn)
        if safe_serialization:
            from safetensors.torch import save_file
            save_file({k: v.contiguous().clone() for k, v in sh.items()}, path, metadata={"format": "pt"})
        else:
            torch.save(sh, path)
        for k in sh:
            index["weight_map"][k] = fn
    if len(shards) > 1:
        with open(os.path.join(output_dir, f"{base}.{ext}.index.json"), "w") as f:
            json.dump(index, f, indent=2, sort_keys=True)
    print(f"Saved fp32 state dict ({len(sd)} tensors, {len(shards)} shard(s)) to {output_dir}")


def load_state_dict_from_zero_checkpoint(model, checkpoint_dir, tag=None):
    sd = get_fp32_state_dict_from_zero_checkpoint(checkpoint_dir, tag)
    model = model.cpu()
    model.load_state_dict(sd, strict=False)
    return model


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("checkpoint_dir", type=str, help="path to the checkpoint folder, e.g. path/checkpoint-12")
    ap.add_argument("output_dir", type=str, help="directory for the consolidated fp32 weights")
    ap.add_argument("--max_shard_size", type=str, default="5GB")
    ap.add_argument("--safe_serialization", default=False, action="store_true")
    ap.add_argument("-t", "--tag", type=str, default=None)
    ap.add_argument("--exclude_frozen_parameters", action="store_true")
    ap.add_argument("-d", "--debug", action="store_true")
    a = ap.parse_args()
    convert_zero_checkpoint_to_fp32_state_dict(a.checkpoint_dir, a.output_dir, a.max_shard_size, a.safe_serialization, a.tag,
                                               a.exclude_frozen_parameters)
