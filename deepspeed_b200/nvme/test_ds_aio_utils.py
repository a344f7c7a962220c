"""Helpers of the I/O benchmark (reference ``nvme/test_ds_aio_utils.py``)."""
import os

BYTES_PER_GB = 1024**3
BYTES_PER_MB = 1024**2
BYTES_PER_KB = 1024
LOG_TIDS = [0]


def task_log(tid, msg, force=False):
    if force or tid in LOG_TIDS:
        print(f"tid {tid}: {msg}")


def task_barrier(barrier, num_parties):
    assert barrier.parties == num_parties
    barrier.wait()
    assert not barrier.broken


def report_results(args, read_op, pool_results):
    """``pool_results``: one ``(bytes, seconds)`` per process.  Prints latency + aggregate bandwidth; returns GB/s."""
    label = "Read" if read_op else "Write"
    if None in pool_results or not pool_results:
        print(f"Failure in one of {args.multi_process} {label} processes")
        return None
    total = sum(b for b, _ in pool_results)
    lat = max(t for _, t in pool_results)
    speed = total / lat / BYTES_PER_GB if lat > 0 else float("inf")
    print(f"{label} Latency = {lat} sec")
    print(f"{label} Speed = {speed} GB/sec")
    return speed


def get_block_size_and_count(io_bytes):
    """Largest power-of-two block (<= 1 GB) that divides ``io_bytes`` evenly, and how many of them."""
    block = BYTES_PER_GB
    while block > 1 and io_bytes % block:
        block //= 2
    return block, io_bytes // block


def refine_integer_value(value):
    """``"4K"`` / ``"16m"`` / ``"1G"`` / ``"123"`` -> int."""
    s = str(value).strip()
    mult = {"K": BYTES_PER_KB, "M": BYTES_PER_MB, "G": BYTES_PER_GB}.get(s[-1].upper())
    return int(float(s[:-1]) * mult) if mult else int(s)


def create_filename(folder, read_op, size, tid):
    return os.path.join(folder, f"_aio_{'read' if read_op else 'write'}_{size}.pt.{tid}")


def create_file(filename, num_bytes):
    block, count = get_block_size_and_count(num_bytes)
    chunk = os.urandom(min(block, 1 << 20))
    with open(filename, "wb") as f:
        left = num_bytes
        while left > 0:
            n = min(left, len(chunk))
            f.write(chunk[:n])
            left -= n
    assert os.path.getsize(filename) == num_bytes
    return filename
