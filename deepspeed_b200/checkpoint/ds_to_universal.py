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
    layout = ms["ds_b200_layout"]
    by_mp = Z.get_optim_shards(ds_dir)
    if len(by_mp) > 1:
        raise NotImplementedError("TP-sharded checkpoints: convert each mp rank separately (tp merge uses "
                                  "universal_checkpoint_info patterns)")
    shards = [Z._load(f)["optimizer_state_dict"] for f in by_mp[0]]
    zero_dir = os.path.join(output_folder, "zero")
    os.makedirs(zero_dir, exist_ok=True)
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--input_folder", type=str, required=True, help="<save_dir>/<tag> of the ZeRO checkpoint")
    ap.add_argument("--output_folder", type=str, required=True)
    ap.add_argument("--num_extract_workers", default=4, type=int)
    ap.add_argument("--num_merge_workers", default=2, type=int)
    ap.add_argument("--keep_temp_folder", action="store_true")
    ap.add_argument("--no_strict", dest="strict", action="store_false")
    ap.add_argument("--inject_missing_state", action="store_true")
    a = ap.parse_args()
    convert_to_universal(a.input_folder, a.output_folder, strict=a.strict, inject_missing_state=a.inject_missing_state)


if __name__ == "__main__":
    main()
