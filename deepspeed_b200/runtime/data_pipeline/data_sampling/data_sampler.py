"""Curriculum-aware batch sampler (reference ``data_sampling/data_sampler.py:36 DeepSpeedDataSampler``).

Each step: every metric's scheduler yields the current difficulty; the admissible sample set is the intersection
over metrics of "samples whose metric <= difficulty" (value-based) or "the easiest p percent" (percentile-based);
one global batch is drawn from it without replacement inside an epoch, then sliced per DP rank and micro-batch.
"""
import os

import numpy as np
import torch

from deepspeed_b200.utils.logging import logger
from .. import constants as C
from ..curriculum_scheduler import CurriculumScheduler
from .indexed_dataset import MMapIndexedDataset


class DeepSpeedDataSampler:

    def __init__(self, data_efficiency_config, one_epoch_total_samples, micro_batch_size, data_parallel_rank,
                 data_parallel_size, data_parallel_group, gradient_accumulation_steps, global_rank, drop_last=True):
        self.data_efficiency_config = data_efficiency_config
        self.one_epoch_total_samples = one_epoch_total_samples
        self.index_dtype = np.int64
        ds = data_efficiency_config[C.DATA_SAMPLING]
        self.total_samples = one_epoch_total_samples * ds.get(C.DATA_SAMPLING_NUM_EPOCHS, 1)
        self.micro_batch_size = micro_batch_size
        self.data_parallel_rank, self.data_parallel_size = data_parallel_rank, data_parallel_size
        self.data_parallel_group = data_parallel_group
        self.micro_batch_times_data_parallel_size = micro_batch_size * data_parallel_size
        self.gradient_accumulation_steps = gradient_accumulation_steps
        self.global_batch_size = self.micro_batch_times_data_parallel_size * gradient_accumulation_steps
        self.global_rank = global_rank
        self.drop_last = drop_last
        self.np_rng = np.random.default_rng(data_efficiency_config.get(C.DATA_EFFICIENCY_SEED, 1234))
        self.state = {}
        self.batch = []
        self.consumed_samples = 0
        cl = ds[C.CURRICULUM_LEARNING]
        self.curriculum_learning_enabled = bool(cl.get(C.CURRICULUM_LEARNING_ENABLED, False))
        self.curriculum_step = 0
        self.curriculum_schedulers, self.difficulty_type, self.metric_values, self.sample_index = {}, {}, {}, {}
        self.current_difficulties = {}
        if self.curriculum_learning_enabled:
            for metric, mc in cl[C.CURRICULUM_LEARNING_METRICS].items():
                self.curriculum_schedulers[metric] = CurriculumScheduler(mc)
                self.difficulty_type[metric] = mc[C.CURRICULUM_LEARNING_DIFFICULTY_TYPE]
                i2m = MMapIndexedDataset(mc[C.CURRICULUM_LEARNING_METRIC_PATH])
                i2s = MMapIndexedDataset(mc[C.CURRICULUM_LEARNING_SAMPLE_PATH])
                self.metric_values[metric] = np.asarray([int(i2m[i][0]) for i in range(len(i2m))])
                self.sample_index[metric] = [np.asarray(i2s[i]) for i in range(len(i2s))]
        self._pool = np.zeros(0, dtype=self.index_dtype)
        self._pool_key = None
        assert self.total_samples > 0 and self.micro_batch_size > 0 and data_parallel_size > 0
        assert data_parallel_rank < data_parallel_size

    def __len__(self):
        return self.total_samples

    def set_custom_curriculum_learning_schedule(self, schedule_func_dict):
        for metric, fn in schedule_func_dict.items():
            if metric in self.curriculum_schedulers:
                self.curriculum_schedulers[metric].set_custom_get_difficulty(fn)

    def get_start_end_idx(self, batch_len=None):
        batch_len = batch_len or self.micro_batch_times_data_parallel_size
        per = batch_len // self.data_parallel_size
        return self.data_parallel_rank * per, (self.data_parallel_rank + 1) * per

    def _admissible(self):
        sets = []
        for metric, sched in self.curriculum_schedulers.items():
            d = self.current_difficulties[metric]
            vals, groups = self.metric_values[metric], self.sample_index[metric]
            if self.difficulty_type[metric] == C.CURRICULUM_LEARNING_VALUE_BASED:
                keep = [g for v, g in zip(vals, groups) if v <= d]
            else:  # percentile: easiest d percent of all samples
                order = np.concatenate(groups) if groups else np.zeros(0, self.index_dtype)
                n = max(1, int(len(order) * min(100, d) / 100.0))
                keep = [order[:n]]
            s = np.concatenate(keep) if keep else np.zeros(0, self.index_dtype)
            sets.append(np.unique(s))
        if not sets:
            return np.arange(self.one_epoch_total_samples, dtype=self.index_dtype)
        out = sets[0]
        for s in sets[1:]:
            out = np.intersect1d(out, s, assume_unique=True)
        return out

    # ---- range queries and the cluster view (reference ``data_sampler.py:133-262``) ----------------------------------
    # The reference materialises, per difficulty step, a *cluster* file of the samples that became admissible and draws
    # every batch from the clusters proportionally to their size.  The default path above keeps one shuffled pool of the
    # admissible set instead; these methods expose the same queries / cluster mechanics over in-memory index arrays.
    def get_sample_based_on_metric_value(self, metric, value_start, value_end):
        """Sample ids whose metric value lies in ``(value_start, value_end]`` (``None`` when empty)."""
        rows = [g for v, g in zip(self.metric_values[metric], self.sample_index[metric]) if value_start < v <= value_end]
        return np.concatenate(rows) if rows else None

    def get_sample_based_on_metric_percentile(self, metric, percentile_start, percentile_end):
        """Sample ids ranked (by metric, easiest first) between the two percentiles of the metric's ``max_difficulty`` scale."""
        order = np.concatenate(self.sample_index[metric]) if self.sample_index[metric] else np.zeros(0, self.index_dtype)
        cl = self.data_efficiency_config[C.DATA_SAMPLING][C.CURRICULUM_LEARNING][C.CURRICULUM_LEARNING_METRICS][metric]
        top = cl.get(C.CURRICULUM_LEARNING_MAX_DIFFICULTY, 100)
        per = len(order) // max(1, top)
        lo, hi = per * percentile_start, (len(order) if percentile_end >= top else per * percentile_end)
        return order[lo:hi] if hi > lo else None

    def get_new_cluster(self, previous_difficulties):
        """Append the samples admitted by the move ``previous_difficulties → current_difficulties`` as a new cluster."""
        if not hasattr(self, "data_clusters"):
            self.data_clusters, self.data_cluster_sizes, self.data_cluster_current_position = [], [], []
        now = self._admissible()
        if previous_difficulties:
            cur, self.current_difficulties = self.current_difficulties, dict(previous_difficulties)
            before = self._admissible()
            self.current_difficulties = cur
            now = np.setdiff1d(now, before, assume_unique=True)
        if now.size == 0:
            return False
        self.data_clusters.append(self.np_rng.permutation(now))
        self.data_cluster_sizes.append(int(now.size))
        self.data_cluster_current_position.append(0)
        return True

    def sample_from_clusters(self):
        """How many samples of the next global batch come from each cluster (multinomial over cluster sizes)."""
        w = np.asarray(self.data_cluster_sizes, dtype=np.float64)
        picks = self.np_rng.choice(len(w), self.global_batch_size, replace=True, p=w / w.sum())
        return np.bincount(picks, minlength=len(w))

    def reshuffle_clusters(self, cidx):
        self.data_clusters[cidx] = self.np_rng.permutation(self.data_clusters[cidx])

    def get_sample_from_cluster(self, cidx, num_samples):
        """Next ``num_samples`` ids of cluster ``cidx``; wraps around (after a reshuffle) when the cluster is exhausted."""
        pos = self.data_cluster_current_position[cidx]
        out = list(self.data_clusters[cidx][pos:pos + num_samples])
        self.data_cluster_current_position[cidx] = pos + num_samples
        while len(out) < num_samples:
            self.reshuffle_clusters(cidx)
            more = num_samples - len(out)
            out += list(self.data_clusters[cidx][:more])
            self.data_cluster_current_position[cidx] = more
        return out

    def get_next_global_batch(self):
        if self.curriculum_learning_enabled:
            self.curriculum_step += 1
            for metric, sched in self.curriculum_schedulers.items():
                self.current_difficulties[metric] = sched.update_difficulty(self.curriculum_step)
            key = tuple(sorted(self.current_difficulties.items()))
            if key != self._pool_key:
                adm = self._admissible()
                if adm.size == 0:
                    raise RuntimeError(f"curriculum difficulties {self.current_difficulties} admit no sample")
                self._adm, self._pool_key = adm, key
                self._pool = self.np_rng.permutation(adm)
            need = self.global_batch_size
            out = []
            while need > 0:
                if self._pool.size == 0:
                    self._pool = self.np_rng.permutation(self._adm)
                take = self._pool[:need]
                self._pool = self._pool[need:]
                out.append(take)
                need -= take.size
            self.batch = np.concatenate(out).tolist()
        else:
            start = self.consumed_samples % self.one_epoch_total_samples
            self.batch = [(start + i) % self.one_epoch_total_samples for i in range(self.global_batch_size)]

    def __iter__(self):
        while self.consumed_samples <= self.total_samples:
            if len(self.batch) == 0:
                self.get_next_global_batch()
            cur = self.batch[:self.micro_batch_times_data_parallel_size]
            self.batch = self.batch[self.micro_batch_times_data_parallel_size:]
            if len(cur) == self.micro_batch_times_data_parallel_size or (len(cur) > 0 and not self.drop_last):
                s, e = self.get_start_end_idx(len(cur))
                yield cur[s:e]
                self.consumed_samples += len(cur)
            if len(cur) == 0:
                break

    def state_dict(self):
        return {C.CURRICULUM_LEARNING_BATCH: self.batch, C.CURRICULUM_LEARNING_CONSUMED_SAMPLES: self.consumed_samples,
                C.CURRICULUM_LEARNING_STEP: self.curriculum_step,
                C.CURRICULUM_LEARNING_CURRENT_DIFFICULTIES: self.current_difficulties,
                C.CURRICULUM_LEARNING_NP_RNG_STATE: self.np_rng.bit_generator.state}

    def load_state_dict(self, sd):
        self.batch = sd[C.CURRICULUM_LEARNING_BATCH]
        self.consumed_samples = sd[C.CURRICULUM_LEARNING_CONSUMED_SAMPLES]
        self.curriculum_step = sd[C.CURRICULUM_LEARNING_STEP]
        self.current_difficulties = sd[C.CURRICULUM_LEARNING_CURRENT_DIFFICULTIES]
        self.np_rng.bit_generator.state = sd[C.CURRICULUM_LEARNING_NP_RNG_STATE]
        for metric, sched in self.curriculum_schedulers.items():
            if metric in self.current_difficulties:
                sched.set_current_difficulty(self.current_difficulties[metric])
        self._pool_key = None
