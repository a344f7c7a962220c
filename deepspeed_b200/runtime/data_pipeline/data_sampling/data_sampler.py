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
