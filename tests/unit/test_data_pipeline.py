import numpy as np
import pytest
import torch
from torch import nn


def test_curriculum_schedules():
    from deepspeed_b200.runtime.data_pipeline.curriculum_scheduler import CurriculumScheduler
    s = CurriculumScheduler({"min_difficulty": 8, "max_difficulty": 64, "schedule_type": "fixed_linear",
                             "schedule_config": {"total_curriculum_step": 100, "difficulty_step": 8}})
    assert s.update_difficulty(0) == 8 and s.update_difficulty(50) == 32 and s.update_difficulty(100) == 64
    assert s.update_difficulty(1000) == 64
    r = CurriculumScheduler({"min_difficulty": 8, "max_difficulty": 64, "schedule_type": "fixed_root",
                             "schedule_config": {"total_curriculum_step": 100, "difficulty_step": 8, "root_degree": 2}})
    assert r.get_difficulty(25) == 32  # sqrt(0.25) = 0.5 -> 8 + 28 = 36 -> floor to 32
    d = CurriculumScheduler({"min_difficulty": 1, "max_difficulty": 3, "schedule_type": "fixed_discrete",
                             "schedule_config": {"difficulty": [1, 2, 3], "max_step": [5, 10]}})
    assert [d.get_difficulty(i) for i in (1, 5, 6, 10, 11)] == [1, 1, 2, 2, 3]
    c = CurriculumScheduler({"min_difficulty": 1, "max_difficulty": 9, "schedule_type": "custom"})
    c.set_custom_get_difficulty(lambda step: min(9, step))
    assert c.update_difficulty(4) == 4


def test_indexed_dataset_analyzer_sampler(tmp_path):
    from deepspeed_b200.runtime.data_pipeline.data_sampling import (DataAnalyzer, DeepSpeedDataSampler, MMapIndexedDataset,
                                                                   MMapIndexedDatasetBuilder)
    b = MMapIndexedDatasetBuilder(str(tmp_path / "d.bin"), dtype=np.int32)
    rows = [np.arange(n, dtype=np.int32) for n in (3, 5, 1, 4)]
    for r in rows:
        b.add_item(r)
        b.end_document()
    b.finalize(str(tmp_path / "d.idx"))
    ds = MMapIndexedDataset(str(tmp_path / "d"))
    assert len(ds) == 4 and all(np.array_equal(ds[i], rows[i]) for i in range(4))
    assert np.array_equal(ds.get(1, 2, 2), rows[1][2:4])

    class Lens(torch.utils.data.Dataset):
        lens = [3, 9, 5, 12, 7, 3, 10, 6]

        def __len__(self):
            return len(self.lens)

        def __getitem__(self, i):
            return torch.tensor(self.lens[i])

    an = DataAnalyzer(Lens(), num_workers=2, worker_id=0, batch_size=3, metric_names=["seqlen"],
                      metric_functions=[lambda batch: batch.numpy()], metric_types=["single_value_per_sample"],
                      metric_dtypes=[np.int64], save_path=str(tmp_path / "an"))
    an.run_map()
    an2 = DataAnalyzer(Lens(), num_workers=2, worker_id=1, batch_size=3, metric_names=["seqlen"],
                       metric_functions=[lambda batch: batch.numpy()], metric_types=["single_value_per_sample"],
                       metric_dtypes=[np.int64], save_path=str(tmp_path / "an"))
    an2.run_map()
    an.run_reduce()
    base = str(tmp_path / "an" / "seqlen" / "seqlen")
    i2m = MMapIndexedDataset(base + "_index_to_metric")
    assert [int(i2m[i][0]) for i in range(len(i2m))] == sorted(set(Lens.lens))
    cfg = {"seed": 1, "data_sampling": {"num_epochs": 2, "curriculum_learning": {"enabled": True, "curriculum_metrics": {
        "seqlen": {"index_to_sample_path": base + "_index_to_sample", "index_to_metric_path": base + "_index_to_metric",
                   "difficulty_type": "value", "clustering_type": "single_cluster", "min_difficulty": 5,
                   "max_difficulty": 12, "schedule_type": "fixed_linear",
                   "schedule_config": {"total_curriculum_step": 4, "difficulty_step": 1}}}}}}
    sampler = DeepSpeedDataSampler(cfg, 8, micro_batch_size=2, data_parallel_rank=0, data_parallel_size=1,
                                   data_parallel_group=None, gradient_accumulation_steps=1, global_rank=0)
    it = iter(sampler)
    first = next(it)
    assert all(Lens.lens[i] <= 5 for i in first)  # step 1: difficulty floor(5 + 7/4) = 6 -> lens <= 6
    for _ in range(4):
        batch = next(it)
    assert len(batch) == 2
    sd = sampler.state_dict()
    sampler.load_state_dict(sd)


def test_random_ltd_wrapper_and_scheduler():
    from deepspeed_b200.runtime.data_pipeline.data_routing import RandomLTDScheduler, convert_to_random_ltd
    from deepspeed_b200.runtime.data_pipeline.data_routing.helper import save_without_random_ltd

    class Block(nn.Module):
        def __init__(self):
            super().__init__()
            self.lin = nn.Linear(8, 8)

        def forward(self, x, attention_mask=None):
            return x + self.lin(x)

    model = nn.Sequential(Block(), Block())
    convert_to_random_ltd(model, Block)
    cfg = {"total_layer_num": 2, "random_ltd_layer_num": 2, "global_batch_size": 4, "model_mask_name": None,
           "micro_batch_size": 2, "hidden_state_order": "batch_seq_dim", "model_type": "decoder",
           "random_ltd_schedule": {"min_value": 4, "max_value": 10, "schedule_type": "fixed_linear",
                                   "schedule_config": {"seq_per_step": 2, "require_steps": 1}}}
    sched = RandomLTDScheduler(cfg)
    for i, m in enumerate(model):
        m.init_config(cfg, sched, i)
    sched.update_seq(0)
    assert sched.get_current_seq() == 4
    x = torch.randn(2, 10, 8, requires_grad=True)
    y = model.train()(x)
    assert y.shape == x.shape
    changed = (y - x).abs().sum(-1) > 0
    assert changed.sum(1).max() <= 8 and changed.sum(1).min() >= 4  # each layer touched 4 tokens
    y.sum().backward()
    assert torch.isfinite(x.grad).all()
    sched.update_seq(3)
    assert sched.get_current_seq() == 10
    assert all(".random_ltd_layer" not in k for k in save_without_random_ltd(model))
