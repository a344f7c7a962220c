import base64
import json

import pytest

from deepspeed_b200.runtime.config import DeepSpeedConfig, DeepSpeedConfigError
from deepspeed_b200.runtime.zero.config import DeepSpeedZeroConfig


def test_batch_triad_all_combinations():
    c = DeepSpeedConfig({"train_batch_size": 32, "train_micro_batch_size_per_gpu": 4})
    assert c.gradient_accumulation_steps == 8
    c = DeepSpeedConfig({"train_batch_size": 32, "gradient_accumulation_steps": 2})
    assert c.train_micro_batch_size_per_gpu == 16
    c = DeepSpeedConfig({"train_micro_batch_size_per_gpu": 3, "gradient_accumulation_steps": 5})
    assert c.train_batch_size == 15
    c = DeepSpeedConfig({"train_batch_size": 7})
    assert (c.train_micro_batch_size_per_gpu, c.gradient_accumulation_steps) == (7, 1)
    with pytest.raises(AssertionError):
        DeepSpeedConfig({"train_batch_size": 32, "train_micro_batch_size_per_gpu": 4, "gradient_accumulation_steps": 3})
    with pytest.raises(DeepSpeedConfigError):
        DeepSpeedConfig({})


def test_sources_json_base64_and_duplicates(tmp_path):
    d = {"train_batch_size": 8, "zero_optimization": {"stage": 2}}
    p = tmp_path / "ds.json"
    p.write_text(json.dumps(d))
    assert DeepSpeedConfig(str(p)).zero_optimization_stage == 2
    enc = base64.urlsafe_b64encode(json.dumps(d).encode()).decode()
    assert DeepSpeedConfig(enc).zero_optimization_stage == 2
    dup = tmp_path / "dup.json"
    dup.write_text('{"train_batch_size": 8, "train_batch_size": 16}')
    with pytest.raises(ValueError):
        DeepSpeedConfig(str(dup))


def test_zero_defaults_and_aliases():
    z = DeepSpeedZeroConfig(stage=3)
    assert z.overlap_comm is True and z.reduce_bucket_size == int(5e8) and z.prefetch_bucket_size == int(5e7)
    assert z.param_persistence_threshold == int(1e5) and z.max_live_parameters == int(1e9)
    assert DeepSpeedZeroConfig(stage=2).overlap_comm is False
    z = DeepSpeedZeroConfig(stage=3, stage3_prefetch_bucket_size=123, stage3_max_live_parameters="auto")
    assert z.prefetch_bucket_size == 123 and z.max_live_parameters == int(1e9)
    with pytest.raises(Exception):
        DeepSpeedZeroConfig(stage=3, not_a_key=1)


def test_deprecated_cpu_offload_maps_to_offload_optimizer():
    z = DeepSpeedZeroConfig(stage=2, cpu_offload=True)
    assert z.offload_optimizer is not None and z.offload_optimizer.device == "cpu"


def test_fp16_bf16_exclusive_and_dynamic_scale_args():
    with pytest.raises(DeepSpeedConfigError):
        DeepSpeedConfig({"train_batch_size": 1, "fp16": {"enabled": True}, "bf16": {"enabled": True}})
    c = DeepSpeedConfig({"train_batch_size": 1, "fp16": {"enabled": True, "initial_scale_power": 10, "hysteresis": 3}})
    a = c.dynamic_loss_scale_args
    assert a["init_scale"] == 1024 and a["delayed_shift"] == 3 and c.loss_scale == 0


def test_optimizer_scheduler_blocks():
    c = DeepSpeedConfig({"train_batch_size": 1, "optimizer": {"type": "AdamW", "params": {"lr": 1e-3}},
                         "scheduler": {"type": "WarmupLR", "params": {"warmup_num_steps": 10}}})
    assert c.optimizer_name == "adamw" and c.optimizer_params["lr"] == 1e-3 and c.scheduler_name == "WarmupLR"


def test_tag_validation_values():
    with pytest.raises(DeepSpeedConfigError):
        DeepSpeedConfig({"train_batch_size": 1, "checkpoint": {"tag_validation": "bogus"}})
    assert DeepSpeedConfig({"train_batch_size": 1, "checkpoint": {"tag_validation": "FAIL"}}).checkpoint_tag_validation_fail
