"""Memory-mapped indexed dataset: ``<prefix>.bin`` holds the concatenated samples, ``<prefix>.idx`` the dtype,
sizes, byte offsets and document boundaries (role of reference ``data_sampling/indexed_dataset.py``, the Megatron
mmap format).  The on-disk header here is this framework's own (magic ``DSB2IDX``)."""
import os
import struct

import numpy as np
import torch

_MAGIC = b"DSB2IDX\x00"
_DTYPES = {1: np.uint8, 2: np.int8, 3: np.int16, 4: np.int32, 5: np.int64, 6: np.float32, 7: np.float64, 8: np.uint16,
           9: np.uint32, 10: np.uint64}
_CODES = {np.dtype(v): k for k, v in _DTYPES.items()}


def index_file_path(prefix):
    return prefix + ".idx"


def data_file_path(prefix):
    return prefix + ".bin"


def best_fitting_dtype(vocab_size=None):
    return np.uint16 if vocab_size is not None and vocab_size < 65500 else np.int32


class MMapIndexedDatasetBuilder:

    def __init__(self, out_file, dtype=np.int64):
        self._path = out_file
        self._f = open(out_file, "wb")
        self._dtype = np.dtype(dtype)
        self._sizes, self._docs = [], [0]

    def add_item(self, tensor):
        arr = np.asarray(tensor.numpy() if torch.is_tensor(tensor) else tensor, dtype=self._dtype)
        self._f.write(arr.tobytes(order="C"))
        self._sizes.append(arr.size)

    def add_items(self, arr_list):
        for a in arr_list:
            self.add_item(a)

    def end_document(self):
        self._docs.append(len(self._sizes))

    def merge_file_(self, another_prefix):
        other = MMapIndexedDataset(another_prefix)
        assert other.dtype == self._dtype
        base = len(self._sizes)
        self._sizes.extend(other.sizes.tolist())
        self._docs.extend((base + other.doc_idx[1:]).tolist())
        with open(data_file_path(another_prefix), "rb") as f:
            while True:
                chunk = f.read(1 << 24)
                if not chunk:
                    break
                self._f.write(chunk)

    def finalize(self, index_file):
        self._f.close()
        sizes = np.asarray(self._sizes, dtype=np.int32)
        ptrs = np.zeros(len(sizes), dtype=np.int64)
        if len(sizes) > 1:
            np.cumsum(sizes[:-1].astype(np.int64) * self._dtype.itemsize, out=ptrs[1:])
        docs = np.asarray(self._docs, dtype=np.int64)
        with open(index_file, "wb") as f:
            f.write(_MAGIC)
            f.write(struct.pack("<QBQQ", 1, _CODES[self._dtype], len(sizes), len(docs)))
            f.write(sizes.tobytes())
            f.write(ptrs.tobytes())
            f.write(docs.tobytes())


class MMapIndexedDataset(torch.utils.data.Dataset):

    def __init__(self, path, skip_warmup=True):
        super().__init__()
        self._path = path
        with open(index_file_path(path), "rb") as f:
            assert f.read(8) == _MAGIC, "not a deepspeed_b200 indexed dataset"
            _, code, n, ndocs = struct.unpack("<QBQQ", f.read(25))
            off = f.tell()
        self._dtype = np.dtype(_DTYPES[code])
        self._idx = np.memmap(index_file_path(path), mode="r", order="C")
        self.sizes = np.frombuffer(self._idx, dtype=np.int32, count=n, offset=off)
        self._ptrs = np.frombuffer(self._idx, dtype=np.int64, count=n, offset=off + self.sizes.nbytes)
        self.doc_idx = np.frombuffer(self._idx, dtype=np.int64, count=ndocs, offset=off + self.sizes.nbytes + self._ptrs.nbytes)
        self._bin = np.memmap(data_file_path(path), mode="r", order="C")

    @property
    def dtype(self):
        return self._dtype

    def __len__(self):
        return len(self.sizes)

    def __getitem__(self, idx):
        if isinstance(idx, (int, np.integer)):
            return np.frombuffer(self._bin, dtype=self._dtype, count=int(self.sizes[idx]), offset=int(self._ptrs[idx]))
        if isinstance(idx, slice):
            start, stop, step = idx.indices(len(self))
            if step != 1:
                raise ValueError("Slices into indexed_dataset must be contiguous")
            return [self[i] for i in range(start, stop)]
        raise TypeError(f"bad index {idx!r}")

    def get(self, idx, offset=0, length=None):
        n = int(self.sizes[idx]) - offset if length is None else length
        return np.frombuffer(self._bin, dtype=self._dtype, count=n, offset=int(self._ptrs[idx]) + offset * self._dtype.itemsize)

    @property
    def supports_prefetch(self):
        return False

    @staticmethod
    def exists(path):
        return os.path.exists(index_file_path(path)) and os.path.exists(data_file_path(path))


def make_builder(out_file, impl="mmap", vocab_size=None, dtype=None):
    return MMapIndexedDatasetBuilder(out_file, dtype=dtype or best_fitting_dtype(vocab_size))


def make_dataset(path, impl="mmap", skip_warmup=True):
    if not MMapIndexedDataset.exists(path):
        print(f"Dataset does not exist: {path}")
        return None
    return MMapIndexedDataset(path, skip_warmup)
