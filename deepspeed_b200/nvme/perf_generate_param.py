"""Turn sweep logs into the optimal ``aio`` config block (reference ``nvme/perf_generate_param.py``)."""
import argparse
import json
import os
import re

from .parse_nvme_stats import READ_SPEED, WRITE_SPEED, get_sorted_results
from .perf_sweep_utils import BENCH_LOG_DIR, READ_LOG_DIR, WRITE_LOG_DIR


def parse_arguments(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--log_dir", type=str, default=BENCH_LOG_DIR, help=f"Folder of aio_perf_sweep.py logs. Default is {BENCH_LOG_DIR}")
    return p.parse_args(argv)


def validate_args(args):
    for d in (os.path.join(args.log_dir, READ_LOG_DIR), os.path.join(args.log_dir, WRITE_LOG_DIR)):
        if not os.path.isdir(d):
            print(f"{d} folder is not existent")
            return False
    return True


def convert_to_param(key):
    """Log-name fields ``(op, single|block, overlap|sequential, t<threads>, p<procs>, d<depth>, bs<size>)`` -> config."""
    f = {"single_submit": "true" if key[1] == "single" else "false", "overlap_events": "true" if key[2] == "overlap" else "false"}
    for field in key[3:]:
        m = re.fullmatch(r"(t|p|d|bs)(\w+)", field)
        if m is None:
            continue
        tag, val = m.groups()
        if tag == "t":
            f["thread_count"] = int(val)
        elif tag == "d":
            f["queue_depth"] = int(val)
        elif tag == "bs":
            mult = {"K": 1024, "M": 1024**2, "G": 1024**3}.get(val[-1].upper())
            f["block_size"] = int(float(val[:-1]) * mult) if mult else int(val)
    return f


def generate_aio_param(read_log_dir, write_log_dir):
    """The configuration with the best combined (read + write) bandwidth among those measured for both."""
    _, read_results = get_sorted_results(read_log_dir, READ_SPEED)[1], get_sorted_results(read_log_dir, READ_SPEED)[0]
    write_results = get_sorted_results(write_log_dir, WRITE_SPEED)[0]
    combined = {}
    for k, v in read_results.items():
        wk = ("write", ) + tuple(k[1:])
        if wk in write_results:
            combined[k[1:]] = v + write_results[wk]
    if not combined:
        raise RuntimeError("no configuration has both a read and a write measurement")
    best = max(combined, key=combined.get)
    param = convert_to_param(("x", ) + tuple(best))
    rk, wk = ("read", ) + tuple(best), ("write", ) + tuple(best)
    print(f"Best performance (GB/sec): read = {read_results[rk]:5.2f}, write = {write_results[wk]:5.2f}")
    print(json.dumps({"aio": param}, indent=3))
    return param


def generate_main(log_dir):
    return generate_aio_param(os.path.join(log_dir, READ_LOG_DIR), os.path.join(log_dir, WRITE_LOG_DIR))


def main(argv=None):
    args = parse_arguments(argv)
    if not validate_args(args):
        raise SystemExit(1)
    generate_main(args.log_dir)


if __name__ == "__main__":
    main()
