"""Helpers for the data analyzer / sampler (reference ``runtime/data_pipeline/data_sampling/utils.py``)."""
import numpy as np

from deepspeed_b200.utils import logger

from .indexed_dataset import MMapIndexedDatasetBuilder


def find_fit_int_dtype(min_value, max_value):
    """Narrowest numpy integer type that holds [min_value, max_value]."""
    kinds = (np.uint8, np.uint16, np.uint32, np.uint64) if min_value >= 0 else (np.int8, np.int16, np.int32, np.int64)
    for k in kinds:
        info = np.iinfo(k)
        if info.min <= min_value and max_value <= info.max:
            return k
    return kinds[-1]


def split_index(start_idx, end_idx, num_partitions):
    edges = np.linspace(start_idx, end_idx, dtype=int, num=num_partitions + 1)
    return [(edges[i], edges[i + 1]) for i in range(num_partitions)]


def split_dataset(dataset, num_workers, worker_id, num_threads):
    worker_splits = split_index(0, len(dataset), num_workers)
    thread_splits = split_index(worker_splits[worker_id][0], worker_splits[worker_id][1], num_threads)
    return worker_splits, thread_splits


def create_mmap_dataset_builder(fname, dtype):
    logger.info(f"Creating mmap dataset builder at {fname}.")
    return MMapIndexedDatasetBuilder(f"{fname}.bin", dtype=dtype)


def close_mmap_dataset_builder(builder, fname):
    builder.end_document()
    builder.finalize(f"{fname}.idx")
    logger.info(f"Finalized mmap dataset builder at {fname}.")
