"""Handle-less (one-shot synchronous) read / write benchmark (reference ``nvme/ds_aio_basic.py``): each repetition opens
its own engine, as a cold-start baseline for the persistent-handle numbers of ``ds_aio_handle``.  The work of one task is
described as a schedule of stages (``pre`` → ``main`` × loops → ``post``), the same shape ``ds_aio_handle`` uses."""
import time

import torch

from .test_ds_aio_utils import create_file, create_filename, report_results, task_log


def pre_basic(args, tid, read_op):
    """Stage 1: resolve the target file (created for reads) and allocate the pinned host buffer."""
    _, folder = args.mapping_list[tid % len(args.mapping_list)]
    filename = create_filename(folder, read_op, args.io_size, tid)
    if read_op:
        create_file(filename, args.io_size)
    buf = torch.empty(args.io_size, dtype=torch.uint8)
    if torch.cuda.is_available():
        buf = buf.pin_memory()
    if not read_op:
        buf.fill_(tid % 251)
    task_log(tid, f"{'read' if read_op else 'write'} target {filename}")
    return {"file": filename, "buffer": buf, "elapsed_sec": 0.0, "num_bytes": args.io_size}


def pre_basic_read(pool_params):
    args, tid = pool_params
    return pre_basic(args, tid, True)


def pre_basic_write(pool_params):
    args, tid = pool_params
    return pre_basic(args, tid, False)


def _one_shot(args, ctxt, read_op):
    from deepspeed_b200.ops.aio import aio_handle
    h = aio_handle(args.block_size, args.queue_depth, args.single_submit, not args.sequential_requests, 1)
    t = time.perf_counter()
    (h.sync_pread if read_op else h.sync_pwrite)(ctxt["buffer"], ctxt["file"])
    ctxt["elapsed_sec"] += time.perf_counter() - t
    return ctxt


def main_basic_read(pool_params):
    args, tid, ctxt = pool_params
    return _one_shot(args, ctxt, True)


def main_basic_write(pool_params):
    args, tid, ctxt = pool_params
    return _one_shot(args, ctxt, False)


def post_basic(pool_params):
    _, _, ctxt = pool_params
    ctxt["buffer"] = None
    return ctxt


def get_schedule(args, read_op):
    return {"pre": pre_basic_read if read_op else pre_basic_write, "main": main_basic_read if read_op else main_basic_write,
            "post": post_basic}


def _task(args, tid, read_op):
    sched = get_schedule(args, read_op)
    ctxt = sched["pre"]((args, tid))
    for _ in range(args.loops):
        ctxt = sched["main"]((args, tid, ctxt))
    sched["post"]((args, tid, ctxt))
    return ctxt["num_bytes"] * args.loops, ctxt["elapsed_sec"]


def aio_basic_multiprocessing(args, read_op):
    from multiprocessing import Pool
    if args.multi_process == 1:
        results = [_task(args, 0, read_op)]
    else:
        with Pool(processes=args.multi_process) as pool:
            results = pool.starmap(_task, [(args, t, read_op) for t in range(args.multi_process)])
    return report_results(args, read_op, results)
