"""Extract metrics from benchmark logs (reference ``nvme/parse_nvme_stats.py``).  Log files are named by their
configuration (``read_single_overlap_t8_p1_d32_bs1M.txt``) and contain ``<Read|Write> Speed = X GB/sec`` /
``... Latency = Y sec`` lines."""
import argparse
import os
import re

READ_SPEED, WRITE_SPEED = "read_speed", "write_speed"
READ_LAT, WRITE_LAT = "read_latency", "write_latency"
PERF_METRICS = [READ_SPEED, WRITE_SPEED]
METRIC_SEARCH = {READ_SPEED: "Read Speed", WRITE_SPEED: "Write Speed", READ_LAT: "Read Latency", WRITE_LAT: "Write Latency"}


def parse_arguments(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--log_dir", type=str, required=True, help="Folder of statistics logs")
    p.add_argument("--metric", type=str, required=True, help=f"Performance metric to report: {PERF_METRICS}")
    return p.parse_args(argv)


def extract_value(key, file):
    """Last ``<key> = <float>`` line of ``file`` (None if absent)."""
    rx = re.compile(re.escape(key) + r"\s*=\s*([0-9.eE+-]+|inf)")
    val = None
    with open(file) as f:
        for line in f:
            m = rx.search(line)
            if m:
                val = float(m.group(1))
    return val


def get_file_key(file):
    """Configuration fields of a log file name as a tuple of strings."""
    return tuple(os.path.splitext(os.path.basename(file))[0].split("_"))


def get_thread_count(file):
    """``..._t<threads>_p<procs>_...`` -> threads * procs (1 when the name carries no such fields)."""
    t = p = 1
    for field in get_file_key(file):
        if re.fullmatch(r"t\d+", field):
            t = int(field[1:])
        elif re.fullmatch(r"p\d+", field):
            p = int(field[1:])
    return t * p


def get_metric(file, metric):
    return extract_value(METRIC_SEARCH[metric], file)


def validate_args(args):
    if args.metric not in METRIC_SEARCH:
        print(f"{args.metric} is not a valid metric: {list(METRIC_SEARCH)}")
        return False
    if not os.path.isdir(args.log_dir):
        print(f"{args.log_dir} folder is not existent")
        return False
    return True


def get_results(log_files, metric):
    out = {}
    for f in log_files:
        v = get_metric(f, metric)
        if v is not None:
            out[get_file_key(f)] = v
    return out


def get_sorted_results(log_dir, metric):
    files = [os.path.join(log_dir, f) for f in os.listdir(log_dir) if os.path.isfile(os.path.join(log_dir, f))]
    res = get_results(files, metric)
    return res, sorted(res)


def main(argv=None):
    args = parse_arguments(argv)
    if not validate_args(args):
        raise SystemExit(1)
    results, keys = get_sorted_results(args.log_dir, args.metric)
    for k in keys:
        print(f"{k} = {results[k]}")


if __name__ == "__main__":
    main()
