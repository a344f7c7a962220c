"""Handle-less (one-shot synchronous) read / write benchmark (reference ``nvme/ds_aio_basic.py``): each repetition opens
its own engine, as a cold-start baseline for the persistent-handle numbers of ``ds_aio_handle``."""
import time

import torch

from .test_ds_aio_utils import create_file, create_filename, report_results


def _task(args, tid, read_op):
    from deepspeed_b200.ops.aio import aio_handle
    _, folder = args.mapping_list[tid % len(args.mapping_list)]
    filename = create_filename(folder, read_op, args.io_size, tid)
    if read_op:
        create_file(filename, args.io_size)
    elapsed = 0.0
    for _ in range(args.loops):
        h = aio_handle(args.block_size, args.queue_depth, args.single_submit, not args.sequential_requests, 1)
        buf = h.new_cpu_locked_tensor(args.io_size, torch.empty(0, dtype=torch.uint8))
        t = time.perf_counter()
        (h.sync_pread if read_op else h.sync_pwrite)(buf, filename)
        elapsed += time.perf_counter() - t
        h.free_cpu_locked_tensor(buf)
    return args.io_size * args.loops, elapsed


def aio_basic_multiprocessing(args, read_op):
    from multiprocessing import Pool
    if args.multi_process == 1:
        results = [_task(args, 0, read_op)]
    else:
        with Pool(processes=args.multi_process) as pool:
            results = pool.starmap(_task, [(args, t, read_op) for t in range(args.multi_process)])
    return report_results(args, read_op, results)
