"""Sparse gradient container for embedding layers (reference ``runtime/sparse_tensor.py:13 SparseTensor``)."""
import torch


class SparseTensor:

    def __init__(self, dense_tensor=None):
        self.orig_dense_tensor = dense_tensor
        self.dtype = self.orig_dense_tensor.dtype if dense_tensor is not None else None
        if dense_tensor is not None:
            if dense_tensor.is_sparse:
                dense_tensor = dense_tensor.coalesce()
                self.indices = dense_tensor.indices().flatten()
                self.values = dense_tensor.values()
            else:
                rows = torch.sum(dense_tensor, dim=1)
                self.indices = rows.nonzero().flatten()
                self.values = dense_tensor[self.indices]
            self.dense_size = list(dense_tensor.size())
        else:
            self.indices = self.values = self.dense_size = None

    def to_coo_tensor(self):
        return torch.sparse_coo_tensor(self.indices.unsqueeze(0), self.values, self.dense_size)

    @staticmethod
    def type():
        return "deepspeed.SparseTensor"

    def to_dense(self):
        it = self.indices.unsqueeze(1)
        full = torch.zeros(self.dense_size, device=self.values.device, dtype=self.values.dtype)
        return full.scatter_add_(0, it.expand_as(self.values), self.values)

    def sparse_size(self):
        idx, val = list(self.indices.size()), list(self.values.size())
        i, v, d = 1, 1, 1
        for s in idx:
            i *= s
        for s in val:
            v *= s
        for s in self.dense_size:
            d *= s
        return i + v, d

    def add(self, b):
        assert self.dense_size == b.dense_size
        self.indices = torch.cat([self.indices, b.indices])
        self.values = torch.cat([self.values, b.values])

    def __str__(self):
        ss, ds = self.sparse_size()
        return f"DeepSpeed.SparseTensor(indices_size={self.indices.size()}, values_size={self.values.size()}, " \
               f"dense_size={self.dense_size}, device={self.values.get_device()}, reduction_factor={ds / ss})"

    def __repr__(self):
        return self.__str__()


def sparse_allreduce(st: SparseTensor, group=None):
    """All-gather (indices, values) of every rank, with padding to the max count; returns the summed SparseTensor
    (reference engine.sparse_allreduce)."""
    from deepspeed_b200 import comm as dist
    world = dist.get_world_size(group)
    n = torch.tensor([st.indices.numel()], dtype=torch.long, device=st.values.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    mx = int(max(s.item() for s in sizes))
    pad_i = torch.zeros(mx, dtype=st.indices.dtype, device=st.indices.device)
    pad_v = torch.zeros(mx, *st.values.shape[1:], dtype=st.values.dtype, device=st.values.device)
    pad_i[:st.indices.numel()] = st.indices
    pad_v[:st.indices.numel()] = st.values
    gi = [torch.zeros_like(pad_i) for _ in range(world)]
    gv = [torch.zeros_like(pad_v) for _ in range(world)]
    dist.all_gather(gi, pad_i, group=group)
    dist.all_gather(gv, pad_v, group=group)
    out = SparseTensor()
    out.dense_size, out.dtype = st.dense_size, st.dtype
    out.indices = torch.cat([g[:int(s.item())] for g, s in zip(gi, sizes)])
    out.values = torch.cat([g[:int(s.item())] for g, s in zip(gv, sizes)]) / world
    return out
