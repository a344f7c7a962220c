"""Handle-based read / write benchmark tasks (reference ``nvme/ds_aio_handle.py``): one aio (or GDS) handle per process,
whole-file transfers timed over ``--loops`` repetitions."""
import time
from multiprocessing import Barrier, Pool

import torch

from .test_ds_aio_utils import create_file, create_filename, report_results, task_barrier, task_log


def _make_handle(args):
    if args.use_gds:
        from deepspeed_b200.ops.gds import gds_handle
        return gds_handle(args.block_size, args.queue_depth, args.single_submit, not args.sequential_requests, args.io_parallel)
    from deepspeed_b200.ops.aio import aio_handle
    return aio_handle(args.block_size, args.queue_depth, args.single_submit, not args.sequential_requests, args.io_parallel)


def pre_handle(args, tid, read_op):
    dev, folder = args.mapping_list[tid % len(args.mapping_list)]
    filename = create_filename(folder, read_op, args.io_size, tid)
    if read_op:
        create_file(filename, args.io_size)
    handle = _make_handle(args)
    if args.gpu:
        buf = torch.empty(args.io_size, dtype=torch.uint8, device=f"cuda:{dev}")
        if args.use_gds:
            handle.pin_device_tensor(buf)
    else:
        buf = handle.new_cpu_locked_tensor(args.io_size, torch.empty(0, dtype=torch.uint8))
    if not read_op:
        buf.fill_(tid % 251)
    task_log(tid, f"created handle + {'device' if args.gpu else 'pinned host'} buffer for {filename}")
    return {"file": filename, "handle": handle, "buffer": buf, "elapsed_sec": 0.0, "num_bytes": args.io_size}


def pre_handle_read(pool_params):
    args, tid = pool_params
    return pre_handle(args, tid, True)


def pre_handle_write(pool_params):
    args, tid = pool_params
    return pre_handle(args, tid, False)


def main_parallel_read(pool_params):
    """Asynchronous submit + wait (the handle's worker threads split the file): ``--io_parallel`` > 1 path."""
    args, tid, ctxt = pool_params
    t = time.perf_counter()
    ctxt["handle"].async_pread(ctxt["buffer"], ctxt["file"])
    ctxt["handle"].wait()
    if args.gpu:
        torch.cuda.synchronize()
    ctxt["elapsed_sec"] += time.perf_counter() - t
    return ctxt


def main_parallel_write(pool_params):
    args, tid, ctxt = pool_params
    t = time.perf_counter()
    ctxt["handle"].async_pwrite(ctxt["buffer"], ctxt["file"])
    ctxt["handle"].wait()
    ctxt["elapsed_sec"] += time.perf_counter() - t
    return ctxt


def main_handle_read(pool_params):
    args, tid, ctxt = pool_params
    t = time.perf_counter()
    ctxt["handle"].pread(ctxt["buffer"], ctxt["file"], args.validate, False, 0)
    if args.gpu:
        torch.cuda.synchronize()
    ctxt["elapsed_sec"] += time.perf_counter() - t
    return ctxt


def main_handle_write(pool_params):
    args, tid, ctxt = pool_params
    t = time.perf_counter()
    ctxt["handle"].pwrite(ctxt["buffer"], ctxt["file"], args.validate, False, 0)
    ctxt["elapsed_sec"] += time.perf_counter() - t
    return ctxt


def post_handle(pool_params):
    _, _, ctxt = pool_params
    if not ctxt["buffer"].is_cuda:
        ctxt["handle"].free_cpu_locked_tensor(ctxt["buffer"])
    ctxt["buffer"] = None
    return ctxt


def _aio_handle_task(args, tid, read_op):
    sched = get_schedule(args, read_op)
    ctxt = sched["pre"]((args, tid))
    for _ in range(args.loops):
        sched["main"]((args, tid, ctxt))
    sched["post"]((args, tid, ctxt))
    return ctxt["num_bytes"] * args.loops, ctxt["elapsed_sec"]


def get_schedule(args, read_op):
    parallel = getattr(args, "io_parallel", 1) and getattr(args, "io_parallel", 1) > 1
    if read_op:
        return {"pre": pre_handle_read, "main": main_parallel_read if parallel else main_handle_read, "post": post_handle}
    return {"pre": pre_handle_write, "main": main_parallel_write if parallel else main_handle_write, "post": post_handle}


def aio_handle_multiprocessing(args, read_op):
    if args.multi_process == 1:
        results = [_aio_handle_task(args, 0, read_op)]
    else:
        with Pool(processes=args.multi_process) as pool:
            results = pool.starmap(_aio_handle_task, [(args, t, read_op) for t in range(args.multi_process)])
    return report_results(args, read_op, results)
