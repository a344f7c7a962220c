import pytest

from deepspeed_b200.elasticity import (compute_elastic_config, ElasticityConfigError, ElasticityError,
                                       ElasticityIncompatibleWorldSize, highly_composite_numbers)

BASE = {"elasticity": {"enabled": True, "max_train_batch_size": 10000, "micro_batch_sizes": [8, 12, 16, 17],
                       "min_gpus": 32, "max_gpus": 1500, "min_time": 20, "version": 0.1}}


def test_hcn_table():
    assert highly_composite_numbers(10000)[:14] == (1, 2, 4, 6, 12, 24, 36, 48, 60, 120, 180, 240, 360, 720)
    assert 720720 in highly_composite_numbers(1_000_000)


def test_basic_10k():
    # same expectation as the reference's test_elastic.py::test_basic_10k
    batch, gpus = compute_elastic_config(BASE, target_deepspeed_version="0.1.0")
    for g in gpus:
        assert batch % g == 0
        assert any((batch // g) % mb == 0 for mb in BASE["elasticity"]["micro_batch_sizes"])
    assert batch == 9792 and len(gpus) == 23


def test_world_size_checks():
    import copy
    batch, gpus, mb = compute_elastic_config(BASE, "0.1.0", world_size=64)
    assert batch == 9792 and mb == 17
    with pytest.raises(ElasticityIncompatibleWorldSize):
        compute_elastic_config(BASE, "0.1.0", world_size=128)
    c = copy.deepcopy(BASE)
    c["elasticity"]["enabled"] = False
    with pytest.raises(ElasticityError):
        compute_elastic_config(c, "0.1.0")
    c = copy.deepcopy(BASE)
    c["elasticity"]["micro_batch_sizes"] = [0, 4]
    with pytest.raises(ElasticityConfigError):
        compute_elastic_config(c, "0.1.0")
    c = copy.deepcopy(BASE)
    c["elasticity"]["version"] = 0.3
    with pytest.raises(ElasticityConfigError):
        compute_elastic_config(c, "0.1.0")


def test_v02_model_parallel():
    c = {"elasticity": {"enabled": True, "max_train_batch_size": 2000, "micro_batch_sizes": [2, 4], "min_gpus": 8,
                        "max_gpus": 64, "version": 0.2, "num_gpus_per_node": 8, "model_parallel_size": 2}}
    batch, gpus, mb = compute_elastic_config(c, "0.1.0", world_size=16)
    assert batch % (16 // 2 * mb) == 0 and mb in (2, 4)


def test_engine_config_uses_elastic_batch():
    from deepspeed_b200.runtime.config import DeepSpeedConfig
    c = {"elasticity": {"enabled": True, "max_train_batch_size": 64, "micro_batch_sizes": [2, 4], "version": 0.1}}
    c["data_parallel_size"] = 4
    cfg = DeepSpeedConfig(c)
    assert cfg.world_size == 4
    assert cfg.train_batch_size == cfg.train_micro_batch_size_per_gpu * cfg.gradient_accumulation_steps * 4
