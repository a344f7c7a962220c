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


def get_val_by_key(d: dict, k):
    """Depth-first lookup of ``k`` anywhere in a nested dict (reference ``autotuning/utils.py:133``)."""
    if k in d:
        return d[k]
    for sub in (v for v in d.values() if isinstance(v, dict)):
        hit = get_val_by_key(sub, k)
        if hit is not None:
            return hit
    return None


def set_val_by_key(d: dict, k, vv):
    """Overwrite every occurrence of ``k`` in a nested dict."""
    stack = [d]
    while stack:
        cur = stack.pop()
        if k in cur:
            cur[k] = vv
        stack.extend(v for v in cur.values() if isinstance(v, dict))


def fetch_hostfile(hostfile_path):
    """``host slots=N`` lines → ordered ``{host: N}``; ``None`` when the file does not exist (reference ``:150``)."""
    import collections
    from deepspeed_b200.utils.logging import logger
    if not os.path.isfile(hostfile_path):
        logger.warning("Unable to find hostfile, will proceed with training with local resources only.")
        return None
    pool = collections.OrderedDict()
    with open(hostfile_path) as fd:
        for raw in fd:
            line = raw.split("#", 1)[0].strip()
            if not line:
                continue
            m = re.fullmatch(r"(\S+)\s+slots=(\d+)", line)
            if m is None:
                raise ValueError(f"Hostfile is not formatted correctly, unable to proceed with training: {raw!r}")
            host, slots = m.group(1), int(m.group(2))
            if host in pool:
                raise ValueError(f"host {host} is already defined in the hostfile")
            pool[host] = slots
    if not pool:
        raise ValueError("Hostfile is empty or not formatted correctly, unable to proceed with training.")
    return pool


def validate_ds_config(config: dict):
    """Reject ZeRO configs the tuner should not launch: offload sections on a stage that cannot use them."""
    z = config.get("zero_optimization") or {}
    stage = z.get("stage")
    if not z or stage in (None, 0, 1):
        return True
    on = lambda key: bool(z.get(key))
    if stage == 2:
        return not (on("cpu_offload") and on("cpu_offload_params"))
    if stage == 3:
        off_p, off_o = z.get("offload_param") or {}, z.get("offload_optimizer") or {}
        nvme = "nvme" in (off_p.get("device"), off_o.get("device"))
        if nvme:
            aio = config.get("aio")
            return bool(aio) and all(os.path.isdir(str(s.get("nvme_path", ""))) for s in (off_p, off_o)
                                     if s.get("device") == "nvme")
        return True
    return True


def remove_dupe_dicts(l):
    """Unique (nested) dicts of ``l``, first occurrence kept."""
    seen, out = set(), []
    for d in l:
        key = json.dumps(d, sort_keys=True)
        if key not in seen:
            seen.add(key)
            out.append(json.loads(key))
    return out


def prune_config(config, ignored_keys=()):
    """Delete the ``ignored_keys`` sections (wherever they are nested) in place."""
    for k in ignored_keys or ():
        del_if_exists(k, config)
    return config


def prune_configs(configs, ignored_keys=()):
    return remove_dupe_dicts([prune_config(c, ignored_keys) for c in configs])


def get_tuning_keys(tuning_space: dict):
    """Names of the parameters with more than one candidate value."""
    keys = []
    for name, val in tuning_space.items():
        if isinstance(val, dict):
            keys += get_tuning_keys(val)
        elif isinstance(val, list) and len(val) > 1:
            keys.append(name)
    return keys
