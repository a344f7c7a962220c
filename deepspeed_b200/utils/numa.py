"""CPU-core / NUMA binding for launched ranks (reference ``utils/numa.py``).

The topology comes from ``/sys/devices/system/node`` (always there on Linux) with ``numactl --hardware`` as a fallback,
so computing a binding does not need the numactl package; only *applying* it (the returned command prefix) does.
"""
import glob
import os
import re
import shutil
import subprocess


def parse_range(rng):
    """``"3"`` -> [3]; ``"2-5"`` -> [2, 3, 4, 5]."""
    m = re.fullmatch(r"\s*(\d+)(?:\s*-\s*(\d+))?\s*", rng)
    if m is None:
        raise ValueError(f"Bad range: '{rng}'")
    lo = int(m.group(1))
    hi = int(m.group(2)) if m.group(2) is not None else lo
    if hi < lo:
        raise ValueError(f"Bad range: '{rng}' (start > end)")
    return list(range(lo, hi + 1))


def parse_range_list(range_str):
    """``"0-3,8,10-11"`` -> sorted unique core ids; overlapping / descending pieces are rejected."""
    out = []
    for piece in range_str.split(","):
        vals = parse_range(piece)
        if out and vals[0] <= out[-1]:
            raise ValueError(f"Bad range list: '{range_str}' (pieces must be ascending and disjoint)")
        out.extend(vals)
    return out


def get_numa_cores():
    """List of core-id lists, one per NUMA node."""
    nodes = sorted(glob.glob("/sys/devices/system/node/node[0-9]*"), key=lambda p: int(re.search(r"(\d+)$", p).group(1)))
    out = []
    for n in nodes:
        try:
            with open(os.path.join(n, "cpulist")) as f:
                txt = f.read().strip()
            out.append(parse_range_list(txt) if txt else [])
        except (OSError, ValueError):
            out = []
            break
    if out:
        return out
    try:
        txt = subprocess.check_output(["numactl", "--hardware"]).decode()
        return [list(map(int, m.group(1).split())) for m in re.finditer(r"node \d+ cpus:([ \d]*)", txt)]
    except (OSError, subprocess.CalledProcessError):
        return [list(range(os.cpu_count() or 1))]


def check_for_numactl_pkg():
    if shutil.which("numactl") is None:
        print("numactl is not found on PATH; install it (apt/yum/pacman package 'numactl') to apply core bindings")
        return False
    return True


def get_numactl_cmd(bind_core_list, num_local_procs, local_rank):
    """Split the usable cores evenly over the local ranks and return ``(cores_per_rank, ["numactl", ...])`` for
    ``local_rank``; memory is bound to the NUMA node(s) that own the chosen cores."""
    check_for_numactl_pkg()
    if "KMP_AFFINITY" in os.environ:
        raise ValueError("Environment variable KMP_AFFINITY conflicts with numactl because it interferes with how many "
                         "CPU cores numactl can set. Unset KMP_AFFINITY before launching with core binding.")
    numa = get_numa_cores()
    cores = parse_range_list(bind_core_list) if bind_core_list else sorted(c for node in numa for c in node)
    per_rank = len(cores) // num_local_procs
    assert per_rank >= 1, "At least one core needs to be available for each rank"
    mine = cores[local_rank * per_rank:(local_rank + 1) * per_rank]
    cmd = ["numactl"]
    owners = [i for i, node in enumerate(numa) if set(mine) & set(node)]
    if owners and all(set(mine) <= set(c for i in owners for c in numa[i]) for _ in (0, )):
        cmd += ["-m", ",".join(map(str, owners))]
    # compress the core list into ranges
    spans, start, prev = [], mine[0], mine[0]
    for c in mine[1:] + [None]:
        if c is None or c != prev + 1:
            spans.append(f"{start}-{prev}" if prev != start else f"{start}")
            start = c
        prev = c if c is not None else prev
    cmd += ["-C", ",".join(spans)]
    return per_rank, cmd
