"""Per-collective timing + bandwidth accounting.

Parity target: reference ``utils/comms_logging.py`` (``calc_bw_log :34``, ``CommsLogger :67``).
Timing here uses CUDA events recorded on the issuing stream (no device-wide synchronize inside
the training loop); the events are resolved lazily when a summary is requested.
"""
import math
from collections import defaultdict

from deepspeed_b200.utils.logging import log_dist


def convert_size(size_bytes):
    if size_bytes == 0:
        return "0B"
    names = ("B", "KB", "MB", "GB", "TB", "PB")
    i = min(int(math.floor(math.log(size_bytes, 1024))), len(names) - 1)
    return f"{round(size_bytes / math.pow(1024, i), 2)} {names[i]}"


def calc_bw_log(comm_op: str, size_bytes: int, duration_ms: float, world: int):
    """Return (message_size, algbw_GBps, busbw_GBps) using the nccl-tests conventions."""
    duration_s = max(duration_ms, 1e-6) / 1e3
    n = max(world, 1)
    if comm_op in ("all_to_all_single", "all_to_all"):
        tput = size_bytes / duration_s
        busbw = tput * ((n - 1) / n)
    elif comm_op in ("all_gather", "all_gather_into_tensor", "reduce_scatter", "reduce_scatter_tensor",
                     "all_gather_coalesced", "reduce_scatter_coalesced"):
        size_bytes = size_bytes * n
        tput = size_bytes / duration_s
        busbw = tput * ((n - 1) / n)
    elif comm_op in ("all_reduce", "all_reduce_coalesced", "inference_all_reduce"):
        tput = size_bytes * 2 / duration_s
        busbw = (size_bytes / duration_s) * (2 * (n - 1) / n)
    else:  # send/recv/broadcast/reduce/gather/scatter/barrier
        tput = size_bytes / duration_s
        busbw = tput
    return size_bytes, tput / 1e9, busbw / 1e9


class CommsLogger:

    def __init__(self):
        self.enabled = False
        self.verbose = False
        self.debug = False
        self.prof_all = True
        self.prof_ops = []
        # op -> msg_size -> [count, [latencies], [algbw], [busbw]]
        self.comms_dict = defaultdict(dict)
        self._pending = []  # (op, size, world, start_evt, end_evt, debug_name)

    def configure(self, cfg):
        self.enabled = cfg.enabled
        self.verbose = cfg.verbose
        self.debug = cfg.debug
        self.prof_all = cfg.prof_all
        self.prof_ops = list(cfg.prof_ops)

    def should_profile(self, op_name, prof_flag=False):
        if not self.enabled:
            return False
        return self.prof_all or prof_flag or op_name in self.prof_ops

    def start_profiling_comms(self):
        self.prof_all = True

    def stop_profiling_comms(self):
        self.prof_all = False

    def start_profiling_op(self, op_name_list):
        self.prof_ops = list(set(self.prof_ops) | set(op_name_list))

    def stop_profiling_op(self, op_name_list):
        self.prof_ops = [o for o in self.prof_ops if o not in op_name_list]

    def defer(self, op, size, world, start_evt, end_evt, name=None):
        self._pending.append((op, size, world, start_evt, end_evt, name))

    def _drain(self):
        for op, size, world, s, e, name in self._pending:
            try:
                e.synchronize()
                ms = s.elapsed_time(e)
            except Exception:
                continue
            self.append(op, name or op, ms, size, world)
        self._pending.clear()

    def append(self, op, record_name, latency_ms, msg_size, world=1):
        size, algbw, busbw = calc_bw_log(op, msg_size, latency_ms, world)
        slot = self.comms_dict[record_name].setdefault(size, [0, [], [], []])
        slot[0] += 1
        slot[1].append(latency_ms)
        slot[2].append(algbw)
        slot[3].append(busbw)
        if self.verbose:
            log_dist(f"comm op: {record_name} | time (ms): {latency_ms:.3f} | msg size: {convert_size(size)} | "
                     f"algbw (GB/s): {algbw:.2f} | busbw (GB/s): {busbw:.2f}", ranks=[0])

    def summary(self):
        self._drain()
        rows = []
        for name, by_size in self.comms_dict.items():
            for size, (count, lats, algs, buss) in sorted(by_size.items()):
                rows.append({
                    "op": name,
                    "size_bytes": size,
                    "count": count,
                    "total_ms": sum(lats),
                    "avg_ms": sum(lats) / len(lats),
                    "algbw_gbps": sum(algs) / len(algs),
                    "busbw_gbps": sum(buss) / len(buss),
                })
        return rows

    def log_all(self, print_log=True, show_straggler=False):
        rows = self.summary()
        if show_straggler:
            import torch
            import torch.distributed as dist
            if dist.is_initialized():
                for r in rows:
                    t = torch.tensor([r["avg_ms"]], dtype=torch.float64)
                    if dist.get_backend() == "nccl":
                        t = t.cuda()
                    tmin = t.clone()
                    dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
                    r["straggler_ms"] = float(r["avg_ms"] - tmin.item())
        if print_log:
            hdr = f"{'Comm. Op':<28}{'Message Size':<16}{'Count':<8}{'Total ms':<12}{'Avg ms':<10}{'algbw GB/s':<12}{'busbw GB/s':<12}"
            lines = [hdr]
            for r in rows:
                lines.append(f"{r['op']:<28}{convert_size(r['size_bytes']):<16}{r['count']:<8}{r['total_ms']:<12.2f}"
                             f"{r['avg_ms']:<10.3f}{r['algbw_gbps']:<12.2f}{r['busbw_gbps']:<12.2f}" +
                             (f"  straggler {r['straggler_ms']:.3f} ms" if "straggler_ms" in r else ""))
            log_dist("\n".join(lines), ranks=[0])
        return rows


def get_caller_func(frame=3):
    """Name of the function ``frame`` levels up the stack (used to label collectives in debug logs)."""
    import sys
    return sys._getframe(frame).f_code.co_name
