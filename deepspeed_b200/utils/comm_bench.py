"""``ds_bench``: collective micro-benchmarks (all_reduce / all_gather / reduce_scatter / all_to_all / broadcast /
pt2pt) over message sizes, NCCL vs this repo's symmetric-memory kernels.  Reference: ``bin/ds_bench`` ->
``benchmarks/communication``.  Device timing with CUDA events, max over ranks (never wall clock)."""
import argparse
import json
import os

import torch


def _bw(op, size_bytes, dur_s, n):
    alg = size_bytes / dur_s / 1e9
    bus = {"all_reduce": 2 * (n - 1) / n, "all_gather": (n - 1) / n, "reduce_scatter": (n - 1) / n,
           "all_to_all": (n - 1) / n}.get(op, 1.0) * alg
    return alg, bus


def run(op, sizes, dtype=torch.bfloat16, trials=20, warmups=5, backend="nccl", use_symm=False):
    import torch.distributed as td
    from deepspeed_b200 import comm as dist
    if not td.is_initialized():
        dist.init_distributed(dist_backend=backend)
    rank, world = td.get_rank(), td.get_world_size()
    cuda = torch.cuda.is_available() and backend == "nccl"
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0))) if cuda else torch.device("cpu")
    if cuda:
        torch.cuda.set_device(dev)
    rows = []
    for nbytes in sizes:
        n = max(world, nbytes // dtype.itemsize // world * world)
        x = torch.ones(n, dtype=dtype, device=dev)
        out = torch.empty(n * world if op == "all_gather" else (n // world if op == "reduce_scatter" else n), dtype=dtype,
                          device=dev)

        def call():
            if op == "all_reduce":
                td.all_reduce(x)
            elif op == "all_gather":
                td.all_gather_into_tensor(out, x)
            elif op == "reduce_scatter":
                td.reduce_scatter_tensor(out, x)
            elif op == "all_to_all":
                dist.all_to_all_single(out, x)
            elif op == "broadcast":
                td.broadcast(x, 0)
            elif op == "pt2pt":
                if rank == 0:
                    td.send(x, 1)
                elif rank == 1:
                    td.recv(x, 0)
            else:
                raise ValueError(op)

        for _ in range(warmups):
            call()
        if cuda:
            torch.cuda.synchronize()
            td.barrier()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(trials):
                call()
            e.record()
            torch.cuda.synchronize()
            dur = s.elapsed_time(e) / 1e3 / trials
        else:
            import time
            td.barrier()
            t = time.perf_counter()
            for _ in range(trials):
                call()
            dur = (time.perf_counter() - t) / trials
        t = torch.tensor([dur], dtype=torch.float64, device=dev)
        td.all_reduce(t, op=td.ReduceOp.MAX)
        dur = float(t.item())
        alg, bus = _bw(op, n * dtype.itemsize, dur, world)
        rows.append({"op": op, "bytes": n * dtype.itemsize, "us": dur * 1e6, "algbw_GBps": alg, "busbw_GBps": bus})
        if rank == 0:
            print(f"{op:>15} {n * dtype.itemsize:>14} B  {dur * 1e6:>10.1f} us  algbw {alg:8.2f} GB/s  busbw {bus:8.2f} GB/s")
    return rows


def main(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--op", default="all", choices=["all", "all_reduce", "all_gather", "reduce_scatter", "all_to_all",
                                                   "broadcast", "pt2pt"])
    p.add_argument("--minsize", type=int, default=1 << 12)
    p.add_argument("--maxsize", type=int, default=1 << 28)
    p.add_argument("--trials", type=int, default=20)
    p.add_argument("--warmups", type=int, default=5)
    p.add_argument("--dtype", default="bfloat16")
    p.add_argument("--backend", default="nccl" if torch.cuda.is_available() else "gloo")
    p.add_argument("--json", default=None)
    p.add_argument("--local_rank", type=int, default=0)
    a = p.parse_args(argv)
    sizes, s = [], a.minsize
    while s <= a.maxsize:
        sizes.append(s)
        s *= 4
    ops = ["all_reduce", "all_gather", "reduce_scatter", "all_to_all", "broadcast"] if a.op == "all" else [a.op]
    out = []
    for op in ops:
        out += run(op, sizes, getattr(torch, a.dtype), a.trials, a.warmups, a.backend)
    if a.json and int(os.environ.get("RANK", 0)) == 0:
        with open(a.json, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
