"""File discovery + state merging helpers for checkpoint reshaping (reference ``checkpoint/reshape_utils.py:1-113``)."""
import os
import re
from collections import OrderedDict

import torch

from .constants import BF16_ZERO_FILE_PREFIX, FP16_ZERO_FILE_PREFIX, MODEL_FILE_PREFIX, ZERO_FILE_PREFIX


def basic_folder_validation(dir):
    assert os.path.exists(dir), f"{dir} path does not exist"
    assert os.path.isdir(dir), f"{dir} is not a folder"


def get_files(dir):
    return [os.path.join(root, f) for root, _, files in os.walk(dir) for f in files]


def get_files_with_prefix(all_files, prefix):
    return sorted(p for p in all_files if os.path.basename(p).startswith(prefix))


def validate_files(file_list):
    missing = [f for f in file_list if not os.path.isfile(f)]
    for f in missing:
        print(f"Error: {f} is not existent")
    return not missing


def sort_zero_files(files, prefix):
    """Order ``<prefix><dp>_mp_rank_<mp>...`` files by (dp rank, mp rank) numerically (not lexically)."""
    rx = re.compile(re.escape(prefix) + r"(\d+)_" + re.escape(MODEL_FILE_PREFIX) + r"(\d+)")
    keyed = []
    for f in files:
        m = rx.search(f)
        if m is None:
            raise ValueError(f"Cannot parse dp_rank and mp_rank from {f}")
        keyed.append((int(m.group(1)), int(m.group(2)), f))
    return [f for _, _, f in sorted(keyed)]


def get_zero_files(dir):
    files = get_files(dir)
    for prefix in (ZERO_FILE_PREFIX, FP16_ZERO_FILE_PREFIX, BF16_ZERO_FILE_PREFIX):
        hits = get_files_with_prefix(files, prefix)
        if hits:
            return sort_zero_files(hits, prefix)
    return []


def partition_data(data_list, num_partitions):
    n = len(data_list)
    assert n % num_partitions == 0, f"{n} items do not split into {num_partitions} equal parts"
    step = n // num_partitions
    return [data_list[i:i + step] for i in range(0, n, step)]


def merge_state(state_a, state_b, key_list=()):
    """Structural merge of two partition states: tensors are concatenated on dim 0, containers recurse, scalars keep ``a``."""
    if type(state_a) is not type(state_b):
        raise ValueError(f"Cannot merge {type(state_a)} with {type(state_b)} at {'.'.join(map(str, key_list))}")
    if isinstance(state_a, (dict, OrderedDict)):
        out = type(state_a)()
        for k, vb in state_b.items():
            out[k] = merge_state(state_a[k], vb, (*key_list, k)) if k in state_a else vb
        return out
    if isinstance(state_a, (list, tuple)):
        if len(state_a) != len(state_b):
            raise ValueError(f"Cannot merge lists of different lengths at {'.'.join(map(str, key_list))}: "
                             f"{len(state_a)} vs {len(state_b)}")
        return type(state_a)(merge_state(a, b, key_list) for a, b in zip(state_a, state_b))
    if torch.is_tensor(state_a):
        return torch.cat([state_a, state_b], 0)
    return state_a


merge_state_dict = lambda a, b, key_list=(): merge_state(a, b, key_list)  # noqa: E731  (reference names)
merge_state_list = lambda a, b, key_list=(): merge_state(list(a), list(b), key_list)  # noqa: E731
