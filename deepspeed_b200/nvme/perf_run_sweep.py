"""Parameter sweep driver (reference ``nvme/perf_run_sweep.py``): every combination of submit mode, overlap, threads,
queue depth and block size is benchmarked for reads and writes, each run logged to a file whose NAME encodes the
configuration (the format ``parse_nvme_stats`` / ``perf_generate_param`` read back)."""
import argparse
import itertools
import json
import os
import sys
import shutil

from .perf import _parse_size, run_io_benchmark
from .perf_sweep_utils import BENCH_LOG_DIR, READ_LOG_DIR, READ_OP_DESC, WRITE_LOG_DIR, WRITE_OP_DESC

DEFAULT_SWEEP_CONFIG = {"block_size": ["128K", "1M"], "queue_depth": [32, 64, 128], "sequential_requests": [True, False],
                        "single_submit": [False], "io_parallel": [1, 2, 8]}


class SweepConfig:

    def __init__(self, args):
        self.folder_to_device_mapping = get_ftd_map(args.nvme_dir)
        self.search_space = get_sweep_config_dict(args.sweep_config)
        self.search_space.update(self.folder_to_device_mapping)
        self.read = not args.no_read
        self.write = not args.no_write
        self.flush_cache = args.flush_page_cache
        self.log_dir = args.log_dir
        self.verbose = args.verbose
        self.other_options = f"--loops {args.loops} --io_size {args.io_size}"
        self.loops, self.io_size, self.use_gds = args.loops, args.io_size, args.gpu and args.use_gds
        self.nvme_dirs = args.nvme_dir


def validate_arguments(args):
    if not args.nvme_dir or not all(os.path.isdir(d) for d in args.nvme_dir):
        print(f"Error: --nvme_dir must list existing folders, got {args.nvme_dir}")
        return False
    if args.no_read and args.no_write:
        print("Error: --no_read and --no_write cannot both be set")
        return False
    if args.use_gds and not args.gpu:
        print("Error: --gpu must be set to transfer with --use_gds")
        return False
    return True


def parse_sweep_arguments(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--nvme_dir", nargs="+", required=True, help="Directory in which to perform I/O tests (NVMe mount points).")
    p.add_argument("--sweep_config", type=str, default=None, help="Performance sweep configuration json file.")
    p.add_argument("--no_read", action="store_true", help="Disable read performance measurements.")
    p.add_argument("--no_write", action="store_true", help="Disable write performance measurements.")
    p.add_argument("--io_size", type=str, default="400M", help="Number of I/O bytes to read/write for performance measurements.")
    p.add_argument("--gpu", action="store_true", help="Test tensor transfers between GPU device and NVME device.")
    p.add_argument("--gds", "--use_gds", dest="use_gds", action="store_true", help="Run the sweep over NVIDIA GPUDirectStorage.")
    p.add_argument("--flush_page_cache", action="store_true",
                   help="Page cache will not be flushed and reported read speeds may be higher than actual ***Requires sudo access***.")
    p.add_argument("--log_dir", type=str, default=BENCH_LOG_DIR, help=f"Output directory for performance log files. Default is {BENCH_LOG_DIR}")
    p.add_argument("--loops", type=int, default=1, help="Count of operation repetitions")
    p.add_argument("--verbose", action="store_true", help="Print debugging information.")
    return p.parse_args(argv)


def dump_cmd_lines(cmd_lines):
    print(f"cmd line count = {len(cmd_lines)}")
    for i, c in enumerate(cmd_lines):
        print(f"{i}: {c}")


def get_ftd_map(nvme_dir_list):
    return {"folder_to_device_mapping": [[f"{d}:{i}" for i, d in enumerate(nvme_dir_list)]]}


def get_sweep_config_dict(sweep_config_json):
    if sweep_config_json is None:
        return dict(DEFAULT_SWEEP_CONFIG)
    with open(sweep_config_json) as f:
        return json.load(f)


def get_sweep_cmd_lines(sweep_config_dict):
    """Cartesian product of the search space as lists of ``--flag value`` tokens."""
    def opts(key, values):
        out = []
        for v in values:
            if isinstance(v, bool):
                out.append([f"--{key}"] if v else [])
            elif isinstance(v, list):
                out.append([f"--{key}"] + [str(x) for x in v])
            else:
                out.append([f"--{key}", str(v)])
        return out
    per_key = [opts(k, v) for k, v in sweep_config_dict.items()]
    return [[tok for part in combo for tok in part] for combo in itertools.product(*per_key)]


def create_cmd_tags(cmd_line):
    tags, i = {}, 0
    while i < len(cmd_line):
        key = cmd_line[i].lstrip("-")
        if i + 1 < len(cmd_line) and not cmd_line[i + 1].startswith("--"):
            tags[key] = cmd_line[i + 1]
            i += 2
        else:
            tags[key] = None
            i += 1
    return tags


def get_log_file(io_op_desc, cmd_line):
    t = create_cmd_tags(cmd_line)
    return "_".join([io_op_desc, "single" if "single_submit" in t else "block", "sequential" if "sequential_requests" in t else "overlap",
                     f"t{t.get('io_parallel', 1)}", f"p{t.get('multi_process', 1)}", f"d{t.get('queue_depth', 32)}",
                     f"bs{t.get('block_size', '1M')}"]) + ".txt"


def remove_folder(folder):
    assert os.path.isdir(folder), f"Error: cannot remove {folder} - folder not found"
    shutil.rmtree(folder)


def flush_page_cache():
    os.system("sync")
    if os.geteuid() == 0:
        with open("/proc/sys/vm/drop_caches", "w") as f:
            f.write("1")


def _run_one(sweep_config, op, cmd_line, log_path):
    t = create_cmd_tags(cmd_line)
    folder = sweep_config.nvme_dirs[0]
    r = run_io_benchmark(os.path.join(folder, f"_aio_bench_{op}.bin"), _parse_size(sweep_config.io_size), read=op == READ_OP_DESC,
                         block_size=_parse_size(t.get("block_size", "1M")), queue_depth=int(t.get("queue_depth", 32)),
                         threads=int(t.get("io_parallel", 1)), single_submit="single_submit" in t,
                         overlap_events="sequential_requests" not in t, loops=sweep_config.loops, use_gds=sweep_config.use_gds)
    label = "Read" if op == READ_OP_DESC else "Write"
    with open(log_path, "w") as f:
        f.write(f"{label} Latency = {r['sec_min']} sec\n{label} Speed = {r['gb_per_s_max']} GB/sec\n")
    return r


def run_sweep_op(sweep_config, op, cmd_lines):
    log_dir = os.path.join(sweep_config.log_dir, READ_LOG_DIR if op == READ_OP_DESC else WRITE_LOG_DIR)
    os.makedirs(log_dir, exist_ok=True)
    out = []
    for c in cmd_lines:
        if sweep_config.flush_cache:
            flush_page_cache()
        if sweep_config.verbose:
            print(op, " ".join(c))
        try:
            out.append(_run_one(sweep_config, op, c, os.path.join(log_dir, get_log_file(op, c))))
        except Exception as e:  # a combination the device rejects is skipped, not fatal
            print(f"skipping {' '.join(c)}: {e}")
    return out


def run_read_sweep(sweep_config, flush_cache_job=None, sync_job=None, cmd_lines=None):
    return run_sweep_op(sweep_config, READ_OP_DESC, cmd_lines)


def run_write_sweep(sweep_config, flush_cache_job=None, sync_job=None, cmd_lines=None):
    return run_sweep_op(sweep_config, WRITE_OP_DESC, cmd_lines)


def script_path():
    return os.path.dirname(os.path.realpath(__file__))


def async_io_setup():
    """Is the native async-io op usable here?"""
    from deepspeed_b200.ops.op_builder import AsyncIOBuilder
    return AsyncIOBuilder().is_compatible()


def gds_io_setup():
    """Is GPUDirect Storage usable here (cuFile present and the GDS op builds)?"""
    from deepspeed_b200.ops.op_builder import GDSBuilder
    return GDSBuilder().is_compatible()


def create_perf_jobs(io_op_desc, log_dir, cmd_lines):
    """One ``Job`` per configuration, each logging to its own file (run with ``launch_sweep``)."""
    from .ds_aio_job import Job
    py = [sys.executable, "-m", "deepspeed_b200.nvme.test_ds_aio", f"--{io_op_desc}"]
    return [Job(cmd_line=py + list(c), output_file=os.path.join(log_dir, get_log_file(io_op_desc, c)))
            for c in cmd_lines]


def launch_sweep(sweep_jobs, sync_job=None, flush_cache_job=None, verbose=False):
    """Run jobs one after another, optionally flushing the page cache / syncing in between."""
    from .ds_aio_job import run_job
    for job in sweep_jobs:
        for aux in (flush_cache_job, sync_job):
            if aux is not None:
                run_job(aux, verbose)
        run_job(job, verbose)


def sweep_main(args):
    sweep_config = SweepConfig(args)
    space = {k: v for k, v in sweep_config.search_space.items() if k != "folder_to_device_mapping"}
    cmd_lines = get_sweep_cmd_lines(space)
    if sweep_config.verbose:
        dump_cmd_lines(cmd_lines)
    os.makedirs(sweep_config.log_dir, exist_ok=True)
    if sweep_config.write:  # writes first: they create the files the reads use
        run_write_sweep(sweep_config, cmd_lines=cmd_lines)
    if sweep_config.read:
        run_read_sweep(sweep_config, cmd_lines=cmd_lines)
    return sweep_config.log_dir


def main(argv=None):
    args = parse_sweep_arguments(argv)
    if not validate_arguments(args):
        raise SystemExit(1)
    log_dir = sweep_main(args)
    from .perf_generate_param import generate_main
    generate_main(log_dir)


if __name__ == "__main__":
    main()
