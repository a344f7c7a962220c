"""Distributed test harness: N processes on one host, file-store rendezvous, gloo (CPU) or NCCL (GPU).

Role parity: reference ``tests/unit/common.py`` (``DistributedTest`` / ``DistributedExec``): same idea
(never multi-node, file:// store in a tmpdir, hard timeout, child exceptions marshalled back), much
smaller: a function-based ``run_distributed(fn, world_size, ...)``.
"""
import os
import sys
import tempfile
import traceback

import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, init_file, fn, args, backend, err_q):
    try:
        os.environ["RANK"] = str(rank)
        os.environ["LOCAL_RANK"] = str(rank)
        os.environ["WORLD_SIZE"] = str(world)
        os.environ["LOCAL_WORLD_SIZE"] = str(world)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", str(29400 + (os.getpid() % 500)))
        os.environ.setdefault("DSB200_LOG_LEVEL", "warning")
        if ROOT not in sys.path:
            sys.path.insert(0, ROOT)
        torch.set_num_threads(1)
        import torch.distributed as dist
        if backend == "nccl":
            torch.cuda.set_device(rank)
        dist.init_process_group(backend, init_method=f"file://{init_file}", rank=rank, world_size=world)
        fn(*args)
        dist.barrier()
        dist.destroy_process_group()
        # the test body is done and the group is torn down: leave without running interpreter / static destructors (under
        # load a c10d helper thread occasionally aborts the exiting process with "terminate called without an active
        # exception", which would be reported as a test failure)
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)
    except Exception:
        err_q.put((rank, traceback.format_exc()))
        raise


def run_distributed(fn, world_size=2, args=(), backend=None, timeout=240):
    """Run ``fn(*args)`` on ``world_size`` ranks; raises AssertionError with the child's traceback."""
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() and torch.cuda.device_count() >= world_size else "gloo"
    ctx = mp.get_context("spawn")
    err_q = ctx.SimpleQueue()
    with tempfile.TemporaryDirectory() as d:
        init_file = os.path.join(d, "rdzv")
        procs = []
        for r in range(world_size):
            p = ctx.Process(target=_worker, args=(r, world_size, init_file, fn, args, backend, err_q), daemon=False)
            p.start()
            procs.append(p)
        import time
        deadline = time.time() + timeout
        failed = None
        while time.time() < deadline:
            alive = [p for p in procs if p.is_alive()]
            bad = [p for p in procs if (not p.is_alive()) and p.exitcode not in (0, None)]
            if bad:
                failed = bad
                break
            if not alive:
                break
            time.sleep(0.05)
        hung = [p for p in procs if p.is_alive()]
        for p in hung:  # our own children only: exact PIDs
            p.terminate()
        for p in procs:
            p.join(timeout=10)
        msgs = []
        while not err_q.empty():
            msgs.append(err_q.get())
        if msgs:
            raise AssertionError("\n".join(f"[rank {r}]\n{tb}" for r, tb in msgs))
        if failed:
            raise AssertionError(f"ranks exited with codes {[p.exitcode for p in failed]}")
        if hung and not failed:
            raise AssertionError(f"distributed test timed out after {timeout}s")
