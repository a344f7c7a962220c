"""Offline data analysis for curriculum learning: compute per-sample metrics (map) and merge them into
``index_to_metric`` / ``index_to_sample`` indexed datasets (reduce).  Reference
``data_sampling/data_analyzer.py`` (``DataAnalyzer :22``, ``DistributedDataAnalyzer :455``)."""
import os
from collections import defaultdict

import numpy as np
import torch

from .indexed_dataset import MMapIndexedDataset, MMapIndexedDatasetBuilder


class DataAnalyzer:

    def __init__(self, dataset, num_workers=1, worker_id=0, num_threads=1, num_threads_reduce=1, specific_threads=(),
                 batch_size=1, metric_names=(), metric_functions=(), metric_types=(), metric_dtypes=(), save_path="./",
                 collate_fn=None, custom_map_init=None, custom_map_update=None, custom_map_finalize=None,
                 custom_reduce=None, sample_indices=None):
        self.dataset = dataset
        self.num_workers, self.worker_id = num_workers, worker_id
        self.batch_size = batch_size
        self.metric_names, self.metric_functions = list(metric_names), list(metric_functions)
        self.metric_types, self.metric_dtypes = list(metric_types), list(metric_dtypes)
        self.save_path = save_path
        self.collate_fn = collate_fn
        self.sample_indices = sample_indices

    def _my_range(self):
        n = len(self.dataset) if self.sample_indices is None else len(self.sample_indices)
        per = (n + self.num_workers - 1) // self.num_workers
        return self.worker_id * per, min(n, (self.worker_id + 1) * per)

    def _wdir(self, name):
        d = os.path.join(self.save_path, name, f"worker{self.worker_id}")
        os.makedirs(d, exist_ok=True)
        return d

    # ---- map: every worker scans its slice and writes per-worker partial results --------------------------------------
    def init_metric_results(self, thread_id=0, metric_names=None, metric_types=None, metric_dtypes=None, save_path=None,
                            worker_id=None):
        """Empty accumulator per metric: ``{"pairs": [(sample, value)...]}`` or ``{"total": None}``."""
        names = metric_names or self.metric_names
        kinds = metric_types or self.metric_types
        return {n: ({"pairs": []} if k == "single_value_per_sample" else {"total": None}) for n, k in zip(names, kinds)}

    def update_metric_results(self, data, metric_types, metric_dtypes, metric_functions, metric_results, batch_start_idx=0,
                              sample_ids=None):
        """Evaluate every metric function on one collated batch and fold the values into ``metric_results``."""
        for name, fn, kind in zip(self.metric_names, metric_functions, metric_types):
            v = fn(data)
            if kind == "single_value_per_sample":
                vals = np.asarray(v).reshape(-1).tolist()
                ids = sample_ids if sample_ids is not None else range(batch_start_idx, batch_start_idx + len(vals))
                metric_results[name]["pairs"].extend(zip(ids, vals))
            elif kind == "accumulate_value_over_samples":
                cur = metric_results[name]["total"]
                metric_results[name]["total"] = v if cur is None else cur + v
            else:
                raise ValueError(f"unknown metric type {kind}")
        return metric_results

    def finalize_metric_results(self, metric_types, metric_dtypes, metric_results):
        """Persist this worker's partial results under ``<save_path>/<metric>/worker<id>/``."""
        for name, kind in zip(self.metric_names, metric_types):
            d = self._wdir(name)
            if kind == "accumulate_value_over_samples":
                np.save(os.path.join(d, "accumulate.npy"), np.asarray(metric_results[name]["total"]))
            else:
                np.save(os.path.join(d, "sample_to_metric.npy"),
                        np.asarray(metric_results[name]["pairs"], dtype=np.int64).reshape(-1, 2))

    def run_map_helper(self, thread_id=0):
        lo, hi = self._my_range()
        results = self.init_metric_results(thread_id)
        for s0 in range(lo, hi, self.batch_size):
            idx = list(range(s0, min(s0 + self.batch_size, hi)))
            if self.sample_indices is not None:
                idx = [self.sample_indices[i] for i in idx]
            batch = [self.dataset[i] for i in idx]
            batch = self.collate_fn(batch) if self.collate_fn else torch.utils.data.default_collate(batch)
            self.update_metric_results(batch, self.metric_types, self.metric_dtypes, self.metric_functions, results,
                                       sample_ids=idx)
        self.finalize_metric_results(self.metric_types, self.metric_dtypes, results)

    def run_map(self):
        self.run_map_helper(0)

    # ---- reduce: worker 0 merges the partial results into the index files the sampler reads --------------------------
    def _builder(self, base, name, suffix, dtype):
        return MMapIndexedDatasetBuilder(os.path.join(base, f"{name}_{suffix}.bin"), dtype=dtype), \
            os.path.join(base, f"{name}_{suffix}.idx")

    def merge_gather_map_stats(self, num_workers, num_threads, num_threads_reduce, t_idx_reduce, metric_save_path, metric_name,
                               return_dict):
        """Concatenate the workers' ``(sample, value)`` tables of one metric (sorted by sample id) into ``return_dict``."""
        pairs = np.concatenate([np.load(os.path.join(metric_save_path, f"worker{w}", "sample_to_metric.npy"))
                                for w in range(num_workers)])
        return_dict[t_idx_reduce] = pairs[np.argsort(pairs[:, 0], kind="stable")]
        return return_dict[t_idx_reduce]

    def merge_sample_to_metric(self, pairs, base, name, dtype):
        """``<metric>_sample_to_metric``: row i holds the metric value of sample i."""
        b, idx = self._builder(base, name, "sample_to_metric", dtype)
        for v in pairs[:, 1]:
            b.add_item(np.asarray([v]))
        b.end_document()
        b.finalize(idx)

    def merge_metric_to_sample(self, pairs, base, name, dtype):
        """``<metric>_index_to_sample`` / ``_index_to_metric``: one row per distinct value (ascending) with its samples."""
        groups = defaultdict(list)
        for i, v in pairs:
            groups[int(v)].append(int(i))
        values = sorted(groups)
        (i2s, i2s_idx), (i2m, i2m_idx) = self._builder(base, name, "index_to_sample", np.int64), \
            self._builder(base, name, "index_to_metric", dtype)
        for v in values:
            i2s.add_item(np.asarray(groups[v], dtype=np.int64))
            i2m.add_item(np.asarray([v]))
        i2s.end_document(), i2m.end_document()
        i2s.finalize(i2s_idx), i2m.finalize(i2m_idx)
        return groups, values

    def get_metric_value_percentiles(self, metric_name, num_sample_per_value, total_num_samples):
        """Log how many samples fall under each difficulty percentile; returns ``{percentile: metric value}``."""
        out, seen, pct = {}, 0, 1
        for v in sorted(num_sample_per_value):
            seen += num_sample_per_value[v]
            while pct <= 100 and seen >= total_num_samples * pct / 100.0:
                out[pct] = v
                pct += 1
        return out

    def output_index_to_sample_percentile(self, groups, values, base, name):
        """``<metric>_index_to_sample_percentile_merged``: samples in difficulty order, cut into 100 equal rows."""
        b, idx = self._builder(base, name, "index_to_sample_percentile_merged", np.int64)
        order = np.concatenate([np.asarray(groups[v], dtype=np.int64) for v in values]) if values else np.zeros(0, np.int64)
        for chunk in np.array_split(order, 100):
            b.add_item(chunk)
        b.end_document()
        b.finalize(idx)

    def merge_map_results(self, dataset, metric_names, metric_types, save_path, num_workers, num_threads, num_threads_reduce):
        for name, kind, dt in zip(metric_names, metric_types, self.metric_dtypes):
            base = os.path.join(save_path, name)
            if kind == "accumulate_value_over_samples":
                total = None
                for w in range(num_workers):
                    a = np.load(os.path.join(base, f"worker{w}", "accumulate.npy"))
                    total = a if total is None else total + a
                b, idx = self._builder(base, name, "metric_value", dt)
                b.add_item(np.asarray(total).reshape(-1))
                b.end_document()
                b.finalize(idx)
                continue
            pairs = self.merge_gather_map_stats(num_workers, num_threads, num_threads_reduce, 0, base, name, {})
            self.merge_sample_to_metric(pairs, base, name, dt)
            groups, values = self.merge_metric_to_sample(pairs, base, name, dt)
            self.output_index_to_sample_percentile(groups, values, base, name)

    def run_reduce(self):
        self.merge_map_results(self.dataset, self.metric_names, self.metric_types, self.save_path, self.num_workers,
                               getattr(self, "num_threads", 1), getattr(self, "num_threads_reduce", 1))

    def run_map_reduce(self, comm_group=None):
        self.run_map()
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            torch.distributed.barrier(group=comm_group)
        if self.worker_id == 0:
            self.run_reduce()
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            torch.distributed.barrier(group=comm_group)


class DistributedDataAnalyzer(DataAnalyzer):
    """One worker per rank of the process group (reference :455): map locally, rank 0 reduces."""

    def __init__(self, dataset, num_workers=1, num_threads=1, worker_id=0, batch_size=1, metric_names=(),
                 metric_functions=(), metric_types=(), save_path="./", collate_fn=None, device="cpu", comm_group=None,
                 sample_indices=None, metric_dtypes=()):
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            num_workers = torch.distributed.get_world_size(comm_group)
            worker_id = torch.distributed.get_rank(comm_group)
        super().__init__(dataset, num_workers=num_workers, worker_id=worker_id, batch_size=batch_size,
                         metric_names=metric_names, metric_functions=metric_functions, metric_types=metric_types,
                         metric_dtypes=metric_dtypes or [np.int64] * len(metric_names), save_path=save_path,
                         collate_fn=collate_fn, sample_indices=sample_indices)
        self.comm_group = comm_group

    def run_map_reduce(self):
        super().run_map_reduce(self.comm_group)


class Dist:
    """Collective helpers over tensors of per-rank different length (reference ``data_analyzer.py:732``)."""

    @staticmethod
    def min_max(tensor, comm_group):
        """Global (min, max); meaningful on rank 0 of the group (reduce, not all-reduce)."""
        from deepspeed_b200 import comm as dist
        lo, hi = tensor.min().clone(), tensor.max().clone()
        dist.reduce(lo, 0, op=dist.ReduceOp.MIN, group=comm_group)
        dist.reduce(hi, 0, op=dist.ReduceOp.MAX, group=comm_group)
        return lo.item(), hi.item()

    @staticmethod
    def gather_v(tensor, dst, comm_group, num_workers, worker_id):
        """Variable-length gather to ``dst``: lengths are exchanged first, payloads padded to the longest, padding dropped
        on arrival.  Returns the list of per-rank tensors on ``dst`` and ``None`` elsewhere."""
        from deepspeed_b200 import comm as dist
        n = torch.tensor([tensor.shape[0]], dtype=torch.int64, device=tensor.device)
        sizes = torch.zeros(num_workers, dtype=torch.int64, device=tensor.device)
        dist.all_gather_into_tensor(sizes, n, group=comm_group)
        longest = int(sizes.max())
        padded = torch.zeros((longest, ) + tuple(tensor.shape[1:]), dtype=tensor.dtype, device=tensor.device)
        padded[:tensor.shape[0]] = tensor
        parts = [torch.empty_like(padded) for _ in range(num_workers)]
        dist.all_gather(parts, padded, group=comm_group)  # gloo/nccl agnostic; the extra copies are tiny index tensors
        if worker_id != dst:
            return None
        return [p[:int(s)] for p, s in zip(parts, sizes)]

    @staticmethod
    def sample_sort(tensor, comm_group, num_workers, n_samples=100):
        """Distributed sample sort of the rows of a 2-D tensor by their first column: every rank ends up with one
        contiguous, locally sorted key range (rank r's keys ≤ rank r+1's)."""
        from deepspeed_b200 import comm as dist
        order = torch.argsort(tensor[:, 0], stable=True)
        tensor = tensor[order]
        if num_workers == 1:
            return tensor
        n = tensor.shape[0]
        if n > 0:
            pick = torch.round(torch.linspace(0, n - 1, n_samples)).long()
            samples = tensor[pick, 0].contiguous()
        else:
            samples = torch.zeros(n_samples, dtype=tensor.dtype, device=tensor.device)
        allsamp = [torch.zeros_like(samples) for _ in range(num_workers)]
        dist.all_gather(allsamp, samples, group=comm_group)
        allsamp = torch.cat(allsamp).sort().values
        cuts = allsamp[torch.round(torch.linspace(0, allsamp.numel() - 1, num_workers + 1)).long()].clone()
        if tensor.is_floating_point():
            cuts[0], cuts[-1] = float("-inf"), float("inf")
        else:
            cuts[0], cuts[-1] = torch.iinfo(tensor.dtype).min, torch.iinfo(tensor.dtype).max
        sends = [tensor[(tensor[:, 0] >= cuts[r]) & (tensor[:, 0] < cuts[r + 1])].contiguous() for r in range(num_workers)]
        me = dist.get_rank(group=comm_group)
        got = []
        for r in range(num_workers):  # one variable-length gather per destination rank
            part = Dist.gather_v(sends[r], r, comm_group, num_workers, me)
            if part is not None:
                got = part
        out = torch.cat(got) if got else tensor[:0]
        return out[torch.argsort(out[:, 0], stable=True)]


def test_compare_both_data_analyzers(dataset, save_path="./_data_analyzer_cmp", num_workers=1, metric_names=("seqlen", ),
                                     metric_functions=None, metric_types=("single_value_per_sample", )):
    """Run the single-process and the distributed analyzer on ``dataset`` and check that they wrote identical index files
    (reference ``data_analyzer.py:843``: a self-check users run before trusting the distributed path)."""
    import filecmp
    import os
    fns = metric_functions or [lambda batch: torch.tensor([len(x) for x in batch]) if isinstance(batch, (list, tuple))
                               else (batch != 0).sum(-1)]
    a_dir, b_dir = os.path.join(save_path, "single"), os.path.join(save_path, "distributed")
    DataAnalyzer(dataset, num_workers=1, worker_id=0, batch_size=4, metric_names=list(metric_names), metric_functions=list(fns),
                 metric_types=list(metric_types), save_path=a_dir).run_map_reduce()
    DistributedDataAnalyzer(dataset, num_workers=num_workers, worker_id=0, batch_size=4, metric_names=list(metric_names),
                            metric_functions=list(fns), metric_types=list(metric_types), save_path=b_dir).run_map_reduce()
    same = True
    for root, _, files in os.walk(a_dir):
        for f in files:
            other = os.path.join(b_dir, os.path.relpath(os.path.join(root, f), a_dir))
            same = same and os.path.exists(other) and filecmp.cmp(os.path.join(root, f), other, shallow=False)
    return same
