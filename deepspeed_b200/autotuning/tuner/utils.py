"""Feature encoding of tuning-space points for the model-based tuner (reference ``autotuning/tuner/utils.py``)."""
import collections.abc
import itertools

import numpy as np

from ..utils import get_list


def index_to_feature(p, dims):
    """Mixed-radix digits of ``p`` (least significant first)."""
    out = []
    for d in dims:
        out.append(p % d)
        p //= d
    return out


def feature_to_index(feature, dims):
    return int(sum(int(np.prod(dims[:j])) * k for j, k in enumerate(feature)))


def dict_to_dims(tuning_space):
    dims = []
    for val in tuning_space.values():
        if isinstance(val, dict):
            dims.extend(dict_to_dims(val))
        else:
            dims.append(len(val) if isinstance(val, list) else 1)
    return dims


def gen_combinations(d: dict):
    keys = list(d.keys())
    choices = (gen_combinations(v) if isinstance(v, dict) else get_list(v) for v in d.values())
    for comb in itertools.product(*choices):
        yield dict(zip(keys, comb))


def flatten(d, parent_key="", sep="_"):
    out = {}
    for k, v in d.items():
        nk = f"{parent_key}{sep}{k}" if parent_key else k
        if isinstance(v, collections.abc.MutableMapping):
            out.update(flatten(v, nk, sep))
        else:
            out[nk] = v
    return out


def dict_to_feature(feature_dict, keys, max_value=None):
    """Numeric feature vector of the entries named in ``keys`` (nested dicts recurse), optionally normalised."""
    feat = []
    for key, val in feature_dict.items():
        if key not in keys or val is None or val == "auto" or key == "autotuning" or val == "":
            continue
        if isinstance(val, dict):
            feat.extend(dict_to_feature(val, keys, None))
        else:
            feat.append(float(val))
    if max_value is not None:
        feat = [f / m if m else f for f, m in zip(feat, max_value)]
    return feat
