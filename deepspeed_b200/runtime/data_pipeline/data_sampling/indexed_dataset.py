"""Memory-mapped indexed dataset: ``<prefix>.bin`` holds the concatenated samples, ``<prefix>.idx`` the dtype,
sizes, byte offsets and document boundaries (role of reference ``data_sampling/indexed_dataset.py``, the Megatron
mmap format).

Three on-disk index layouts are understood:

* ``DSB2IDX`` — this framework's own header (default for new files);
* ``MMIDIDX`` — the Megatron / DeepSpeed mmap layout, read transparently and written with ``fmt="megatron"`` so corpora
  tokenised for the reference load unchanged (and vice versa);
* ``TNTIDX`` — the legacy "lazy"/"cached" layout (``IndexedDataset`` / ``IndexedCachedDataset`` / ``IndexedDatasetBuilder``)."""
import os
import struct

import numpy as np
import torch

_MAGIC = b"DSB2IDX\x00"
_MEGATRON_MAGIC = b"MMIDIDX\x00\x00"
_LEGACY_MAGIC = b"TNTIDX\x00\x00"
# dtype codes of the Megatron/DeepSpeed and legacy layouts (reference ``indexed_dataset.py:101``)
dtypes = {1: (np.uint8, torch.uint8), 2: (np.int8, torch.int8), 3: (np.int16, torch.int16), 4: (np.int32, torch.int32),
          5: (np.int64, torch.int64), 6: (np.uint16, None), 7: (np.uint32, None), 8: (np.uint64, None)}
valid_dtypes = {d for pair in dtypes.values() for d in pair if d is not None}


def code(dtype):
    """Interop dtype code of a numpy / torch dtype."""
    for c, pair in dtypes.items():
        if any(d is not None and (dtype is d or (not isinstance(dtype, torch.dtype) and not isinstance(d, torch.dtype)
                                                    and np.dtype(dtype) == np.dtype(d))) for d in pair):
            return c
    raise ValueError(f"{dtype} not supported. Supported types: {valid_dtypes}")


def read_longs(f, n):
    a = np.empty(n, dtype=np.int64)
    f.readinto(a)
    return a


def write_longs(f, a):
    f.write(np.asarray(a, dtype=np.int64).tobytes())


def get_available_dataset_impl():
    return ["lazy", "cached", "mmap"]


def _magic_of(path):
    with open(index_file_path(path), "rb") as f:
        return f.read(9)


def infer_dataset_impl(path):
    """``"mmap"`` / ``"cached"`` (legacy layout) / ``None`` from the index header."""
    if not os.path.exists(index_file_path(path)):
        print(f"Dataset does not exist: {path}")
        return None
    magic = _magic_of(path)
    if magic[:8] == _MAGIC or magic == _MEGATRON_MAGIC:
        return "mmap"
    if magic[:8] == _LEGACY_MAGIC:
        return "cached"
    return None


def dataset_exists(path, impl):
    return MMapIndexedDataset.exists(path) if impl == "mmap" else IndexedDataset.exists(path)


def create_doc_idx(sizes):
    """Document boundaries of a legacy corpus where an empty sample terminates a document."""
    return [0] + [i + 1 for i, n in enumerate(sizes) if n == 0]


def exscan_from_cumsum_(arr):
    """In place: inclusive scan → exclusive scan."""
    if arr.size:
        arr[1:] = arr[:-1].copy()
        arr[0] = 0


def get_pointers_with_total(sizes, elemsize, dtype):
    """(byte offset of every sample, total bytes)."""
    ptrs = np.asarray(sizes, dtype=dtype) * dtype(elemsize) if len(sizes) else np.zeros(0, dtype=dtype)
    np.cumsum(ptrs, out=ptrs)
    total = int(ptrs[-1]) if ptrs.size else 0
    exscan_from_cumsum_(ptrs)
    return ptrs, total
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

    def __init__(self, out_file, dtype=np.int64, fmt="native"):
        assert fmt in ("native", "megatron")
        self._path = out_file
        self._f = open(out_file, "wb")
        self._dtype = np.dtype(dtype)
        self._fmt = fmt
        self._sizes, self._docs = [], [0]

    def add_item(self, tensor):
        arr = np.asarray(tensor.numpy() if torch.is_tensor(tensor) else tensor, dtype=self._dtype)
        self._f.write(arr.tobytes(order="C"))
        self._sizes.append(arr.size)

    def add_item_numpy(self, np_array):
        """``add_item`` for an array that is already numpy (cast to the dataset dtype if needed)."""
        arr = np_array if np_array.dtype == self._dtype else np_array.astype(self._dtype)
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
            if self._fmt == "megatron":
                f.write(_MEGATRON_MAGIC)
                f.write(struct.pack("<QBQQ", 1, code(self._dtype.type), len(sizes), len(docs)))
            else:
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
            magic = f.read(8)
            if magic == _MAGIC:
                table = _DTYPES
            else:
                assert magic + f.read(1) == _MEGATRON_MAGIC, "index file is neither a DSB2IDX nor an MMIDIDX mmap index"
                table = {c: pair[0] for c, pair in dtypes.items()}
            ver, dcode, n, ndocs = struct.unpack("<QBQQ", f.read(25))
            assert ver == 1
            off = f.tell()
        self._dtype = np.dtype(table[dcode])
        self._idx = np.memmap(index_file_path(path), mode="r", order="C")
        self._sizes = np.frombuffer(self._idx, dtype=np.int32, count=n, offset=off)
        self._ptrs = np.frombuffer(self._idx, dtype=np.int64, count=n, offset=off + self.sizes.nbytes)
        self._doc_idx = np.frombuffer(self._idx, dtype=np.int64, count=ndocs, offset=off + self.sizes.nbytes + self._ptrs.nbytes)
        self._bin = np.memmap(data_file_path(path), mode="r", order="C")

    @property
    def dtype(self):
        return self._dtype

    @property
    def sizes(self):
        return self._sizes

    def size(self, index):
        return self._sizes[index]

    @property
    def doc_idx(self):
        return self._doc_idx

    def get_doc_idx(self):
        return self._doc_idx

    def set_doc_idx(self, doc_idx_):
        """Replace the document boundaries (used when a corpus is re-segmented without rewriting the data file)."""
        self._doc_idx = np.asarray(doc_idx_, dtype=np.int64)

    def __len__(self):
        return len(self._sizes)

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


class IndexedDataset(torch.utils.data.Dataset):
    """Legacy ``TNTIDX`` reader: the index lives in memory, samples are read from the ``.bin`` file on demand
    (reference ``indexed_dataset.py:139``)."""
    _HDR_MAGIC = _LEGACY_MAGIC

    def __init__(self, path):
        super().__init__()
        self.path, self.data_file = path, None
        self.read_index(path)

    def read_index(self, path):
        with open(index_file_path(path), "rb") as f:
            assert f.read(8) == self._HDR_MAGIC, ("Index file doesn't match expected format. Make sure that "
                                                  "--dataset-impl is configured properly.")
            ver, dcode, self.element_size, self._len, self.s, self.doc_count = struct.unpack("<6Q", f.read(48))
            assert ver == 1
            self.dtype = dtypes[dcode][0]
            self.dim_offsets = read_longs(f, self._len + 1)
            self.data_offsets = read_longs(f, self._len + 1)
            self.sizes = read_longs(f, self.s)
            self.doc_idx = read_longs(f, self.doc_count)

    def read_data(self, path):
        self.data_file = open(data_file_path(path), "rb", buffering=0)

    def check_index(self, i):
        if not 0 <= i < self._len:
            raise IndexError("index out of range")

    def __del__(self):
        if getattr(self, "data_file", None):
            self.data_file.close()

    def _shape(self, i):
        return tuple(int(x) for x in self.sizes[self.dim_offsets[i]:self.dim_offsets[i + 1]])

    def _read(self, first, shape):
        if not self.data_file:
            self.read_data(self.path)
        out = np.empty(shape, dtype=self.dtype)
        self.data_file.seek(int(self.data_offsets[first]) * self.element_size)
        self.data_file.readinto(out)
        return out

    def __getitem__(self, idx):
        if isinstance(idx, (int, np.integer)):
            self.check_index(idx)
            return self._read(idx, self._shape(idx))
        start, stop, step = idx.indices(len(self))
        if step != 1:
            raise ValueError("Slices into indexed_dataset must be contiguous")
        lens = [int(np.prod(self._shape(i))) for i in range(start, stop)]
        flat = self._read(start, (sum(lens), )) if lens else np.empty(0, dtype=self.dtype)
        return np.split(flat, np.cumsum(lens)[:-1]) if lens else []

    def __len__(self):
        return self._len

    def num_tokens(self, index):
        return self.sizes[index]

    def size(self, index):
        return self.sizes[index]

    @staticmethod
    def exists(path):
        return os.path.exists(index_file_path(path)) and os.path.exists(data_file_path(path))

    @property
    def supports_prefetch(self):
        return False


class IndexedCachedDataset(IndexedDataset):
    """Legacy reader with an explicit ``prefetch(indices)`` stage that pulls the requested samples into one host buffer."""

    def __init__(self, path):
        super().__init__(path)
        self.cache, self.cache_index = None, {}

    @property
    def supports_prefetch(self):
        return True

    def prefetch(self, indices):
        if all(i in self.cache_index for i in indices):
            return
        want = sorted(set(int(i) for i in indices))
        lens = [int(self.data_offsets[i + 1] - self.data_offsets[i]) for i in want]
        self.cache = np.empty(sum(lens), dtype=self.dtype)
        self.cache_index = {}
        pos = 0
        for i, n in zip(want, lens):
            self.cache_index[i] = pos
            self.cache[pos:pos + n] = self._read(i, (n, ))
            pos += n
        if self.data_file:
            self.data_file.close()
            self.data_file = None

    def __getitem__(self, idx):
        if isinstance(idx, (int, np.integer)):
            self.check_index(idx)
            shape = self._shape(idx)
            at = self.cache_index[int(idx)]
            return self.cache[at:at + int(np.prod(shape))].reshape(shape).copy()
        return [self[i] for i in range(*idx.indices(len(self)))]


class IndexedDatasetBuilder:
    """Writer of the legacy ``TNTIDX`` layout (reference ``indexed_dataset.py:272``)."""

    def __init__(self, out_file, dtype=np.int32):
        self.out_file = open(out_file, "wb")
        self.dtype = dtype
        self.element_size = np.dtype(dtype).itemsize
        self.data_offsets, self.dim_offsets, self.sizes, self.doc_idx = [0], [0], [], [0]

    def add_item(self, tensor):
        arr = np.asarray(tensor.numpy() if torch.is_tensor(tensor) else tensor, dtype=self.dtype)
        self.out_file.write(arr.tobytes(order="C"))
        self.data_offsets.append(self.data_offsets[-1] + arr.size)
        self.sizes.extend(arr.shape)
        self.dim_offsets.append(self.dim_offsets[-1] + arr.ndim)

    def end_document(self):
        self.doc_idx.append(len(self.sizes))

    def merge_file_(self, another_file):
        other = IndexedDataset(another_file)
        assert np.dtype(other.dtype) == np.dtype(self.dtype)
        doc_base, data_base, dim_base = len(self.sizes), self.data_offsets[-1], self.dim_offsets[-1]
        self.data_offsets.extend(int(data_base + o) for o in other.data_offsets[1:])
        self.dim_offsets.extend(int(dim_base + o) for o in other.dim_offsets[1:])
        self.sizes.extend(int(x) for x in other.sizes)
        self.doc_idx.extend(int(doc_base + d) for d in other.doc_idx[1:])
        with open(data_file_path(another_file), "rb") as f:
            for chunk in iter(lambda: f.read(1 << 24), b""):
                self.out_file.write(chunk)

    def finalize(self, index_file):
        self.out_file.close()
        with open(index_file, "wb") as f:
            f.write(_LEGACY_MAGIC)
            f.write(struct.pack("<6Q", 1, code(self.dtype), self.element_size, len(self.data_offsets) - 1, len(self.sizes),
                                len(self.doc_idx)))
            for arr in (self.dim_offsets, self.data_offsets, self.sizes, self.doc_idx):
                write_longs(f, arr)


def make_builder(out_file, impl="mmap", vocab_size=None, dtype=None, fmt="native"):
    if impl == "mmap":
        return MMapIndexedDatasetBuilder(out_file, dtype=dtype or best_fitting_dtype(vocab_size), fmt=fmt)
    return IndexedDatasetBuilder(out_file, **({"dtype": dtype} if dtype is not None else {}))


def make_dataset(path, impl="mmap", skip_warmup=True):
    if not IndexedDataset.exists(path):
        print(f"Dataset does not exist: {path}")
        print("Path should be a basename that both .idx and .bin can be appended to get full filenames.")
        return None
    if impl == "infer":
        impl = infer_dataset_impl(path)
    if impl == "lazy":
        return IndexedDataset(path)
    if impl == "cached":
        return IndexedCachedDataset(path)
    if impl == "mmap":
        return MMapIndexedDataset(path, skip_warmup)
    print(f"Unknown dataset implementation: {impl}")
    return None
