"""Experiment-space helpers (reference ``autotuning/utils.py``)."""
import copy
import itertools
import json
import os
import re


def search_error(filename):
    if not os.path.exists(filename):
        return "stderr.log does not exist"
    with open(filename) as f:
        for line in f:
            for s in ("Error", "error", "ERROR"):
                i = line.find(s)
                if i != -1:
                    return line[i:].strip()
    return None


def was_interruptted(filename):
    if not os.path.exists(filename):
        return "stderr.log does not exist"
    with open(filename) as f:
        return any(re.search(r"KeyboardInterrupt", l) for l in f)


def find_replace_str(value, replace_dict):
    if not isinstance(value, str):
        return str(value)
    for m in re.findall(r"\$[\w]+", value):
        var = m[1:]
        value = value.replace(m, str(replace_dict[var])) if var in replace_dict else value
    return value


def find_replace(target, replace_dict):
    if isinstance(target, dict):
        for k, v in target.items():
            if isinstance(v, str):
                target[k] = find_replace_str(v, replace_dict)
            elif isinstance(v, list):
                target[k] = [find_replace_str(x, replace_dict) for x in v]
            elif isinstance(v, dict):
                find_replace(v, replace_dict)


def get_list(val):
    return val if isinstance(val, list) else [val]


def combine_dict(d, u):
    for k, v in u.items():
        if isinstance(v, dict):
            d[k] = combine_dict(d.get(k, {}), v)
        else:
            if k not in d:
                d[k] = v
            else:
                d[k] = get_list(d[k]) + get_list(v)
    return d


def del_if_exists(t, d):
    if t in d:
        del d[t]
        return
    for v in d.values():
        if isinstance(v, dict):
            del_if_exists(t, v)


def replace_dict(d, u, ignored_keys=()):
    if u is not None:
        for k, v in u.items():
            if k in ignored_keys:
                continue
            if v is None:
                del_if_exists(k, d)
            elif isinstance(v, dict):
                d[k] = replace_dict(d.get(k, {}), v, ignored_keys)
            else:
                d[k] = v
    return d


def flatten(d, parent_key="", sep="_"):
    items = []
    for k, v in d.items():
        nk = f"{parent_key}{sep}{k}" if parent_key else k
        if isinstance(v, dict):
            items.extend(flatten(v, nk, sep).items())
        else:
            items.append((nk, v))
    return dict(items)


def get_all_configs(tuning_space: dict, ignore_keys=None):
    """Cartesian product over every list-valued leaf."""
    def walk(node):
        if isinstance(node, dict):
            keys = [k for k in node if not (ignore_keys and k in ignore_keys)]
            subs = [walk(node[k]) for k in keys]
            out = []
            for combo in itertools.product(*subs):
                d = {k: c for k, c in zip(keys, combo)}
                for k in node:
                    if ignore_keys and k in ignore_keys:
                        d[k] = node[k]
                out.append(d)
            return out
        if isinstance(node, list):
            return list(node)
        return [node]

    return [copy.deepcopy(c) for c in walk(tuning_space)]


def canonical_name(config: dict, tuning_keys=None, prefix="", omit_val=False):
    flat = flatten(config)
    parts = []
    for k in sorted(flat):
        if tuning_keys and not any(k.endswith(t) or t in k for t in tuning_keys):
            continue
        short = "".join(w[0] for w in k.split("_") if w)
        parts.append(short if omit_val else f"{short}{flat[k]}")
    name = "_".join(parts)
    return f"{prefix}_{name}" if prefix else name


def get_first_config(config: dict):
    cfg = copy.deepcopy(config)
    for k, v in cfg.items():
        if isinstance(v, dict):
            cfg[k] = get_first_config(v)
        elif isinstance(v, list):
            cfg[k] = v[0]
    return cfg


def write_experiments(exps: list, exps_dir: str):
    os.makedirs(exps_dir, exist_ok=True)
    paths = []
    for e in exps:
        p = os.path.join(exps_dir, f"{e['name']}.json")
        with open(p, "w") as f:
            json.dump(e, f)
        paths.append(p)
    return paths


def memory_to_string(n, postfix="", units=None, precision=2):
    for u, s in (("T", 1 << 40), ("G", 1 << 30), ("M", 1 << 20), ("K", 1 << 10)):
        if units == u or (units is None and n >= s):
            return f"{round(n / s, precision)} {u}{postfix}"
    return f"{n} {postfix}"


def number_to_string(n, postfix="", units=None, precision=2):
    for u, s in (("B", 1e9), ("M", 1e6), ("K", 1e3)):
        if units == u or (units is None and n >= s):
            return f"{round(n / s, precision)} {u}{postfix}"
    return f"{n} {postfix}"
