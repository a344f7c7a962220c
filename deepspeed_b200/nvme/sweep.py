import argparse
import itertools
import json
import os

from .perf import _parse_size, run_io_benchmark

DEFAULT_SWEEP = {"block_size": ["128K", "1M", "8M"], "queue_depth": [8, 32, 128], "threads": [1, 4, 8],
                 "single_submit": [False], "overlap_events": [True, False]}


def run_sweep(folder, io_size="256M", sweep=None, loops=2, use_gds=False, log_dir=None):
    sweep = sweep or DEFAULT_SWEEP
    keys = list(sweep)
    results = []
    path = os.path.join(folder, "ds_nvme_tune.bin")
    os.makedirs(folder, exist_ok=True)
    for combo in itertools.product(*(sweep[k] for k in keys)):
        kw = dict(zip(keys, combo))
        kw["block_size"] = _parse_size(kw["block_size"])
        for read in (False, True):
            try:
                r = run_io_benchmark(path, _parse_size(io_size), read=read, loops=loops, use_gds=use_gds, **kw)
                results.append(r)
            except Exception as e:  # a combination the device rejects is simply skipped
                results.append({"op": "read" if read else "write", **kw, "error": str(e)})
    if log_dir:
        os.makedirs(log_dir, exist_ok=True)
        with open(os.path.join(log_dir, "sweep_results.json"), "w") as f:
            json.dump(results, f, indent=1)
    return results


def generate_aio_param(results):
    """Pick the config maximising min(read, write) bandwidth -> the ``aio`` block for ds_config."""
    by_cfg = {}
    for r in results:
        if "error" in r:
            continue
        k = (r["block_size"], r["queue_depth"], r["threads"], r["single_submit"], r["overlap_events"])
        by_cfg.setdefault(k, {})[r["op"]] = r["gb_per_s_max"]
    best, best_v = None, -1.0
    for k, v in by_cfg.items():
        if "read" in v and "write" in v and min(v["read"], v["write"]) > best_v:
            best, best_v = (k, v), min(v["read"], v["write"])
    if best is None:
        return None
    (bs, qd, th, ss, oe), v = best
    return {"aio": {"block_size": bs, "queue_depth": qd, "intra_op_parallelism": th, "single_submit": ss,
                    "overlap_events": oe}, "read_GBps": v["read"], "write_GBps": v["write"]}


def sweep_main(argv=None):
    p = argparse.ArgumentParser(description="Sweep aio parameters on an NVMe folder and print the best ds_config block")
    p.add_argument("--nvme_dir", type=str, required=True)
    p.add_argument("--io_size", type=str, default="256M")
    p.add_argument("--sweep_config", type=str, default=None)
    p.add_argument("--log_dir", type=str, default="_aio_bench_logs")
    p.add_argument("--loops", type=int, default=2)
    p.add_argument("--use_gds", action="store_true")
    a = p.parse_args(argv)
    sweep = None
    if a.sweep_config:
        with open(a.sweep_config) as f:
            sweep = json.load(f)
    res = run_sweep(a.nvme_dir, a.io_size, sweep, a.loops, a.use_gds, a.log_dir)
    best = generate_aio_param(res)
    print(json.dumps(best, indent=2))
    return best
