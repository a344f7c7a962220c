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


def test_elastic_agent_supervision_policy():
    """The run-loop policy of DSElasticAgent (reference elastic_agent.py:127-188) as a pure function."""
    from deepspeed_b200.elasticity.elastic_agent import (CONTINUE, FAIL, FINISH, RESCALE, RESTART, supervise_decision)
    assert supervise_decision("SUCCEEDED", 3, 4, 4, 0) == FINISH
    assert supervise_decision("HEALTHY", 3, 4, 4, 0) == CONTINUE
    assert supervise_decision("HEALTHY", 0, 4, 4, 2) == RESCALE          # growth is free: no restart budget needed
    assert supervise_decision("FAILED", 2, 4, 4, 0) == RESTART
    assert supervise_decision("UNHEALTHY", 0, 4, 4, 0) == FAIL
    assert supervise_decision("HEALTHY", 1, 4, 3, 0) == RESTART          # a participant left the rendezvous
    assert supervise_decision("HEALTHY", 0, 4, 4, 0, dead_nodes=1) == FAIL  # lost heartbeat, budget exhausted
    import pytest
    with pytest.raises(RuntimeError):
        supervise_decision("INIT", 1, 1, 1, 0)


def test_elastic_agent_run_loop_restarts_then_finishes():
    """Drive ``DSElasticAgent._invoke_run`` with a scripted worker group: FAILED -> restart -> new node -> rescale ->
    SUCCEEDED; restart budget is charged only for the failure."""
    import types
    from torch.distributed.elastic.agent.server.api import RunResult, WorkerState
    from deepspeed_b200.elasticity.elastic_agent import DSElasticAgent
    agent = DSElasticAgent.__new__(DSElasticAgent)
    script = [WorkerState.FAILED, WorkerState.HEALTHY, WorkerState.HEALTHY, WorkerState.SUCCEEDED]
    waiting = [0, 1, 0, 0]
    calls = {"restart": 0, "barrier": 0, "tick": 0}
    rdzv = types.SimpleNamespace(
        _state_holder=types.SimpleNamespace(state=types.SimpleNamespace(participants={"a": 0, "b": 1}, last_heartbeats={})),
        _settings=None, num_nodes_waiting=lambda: waiting[min(calls["tick"] - 1, 3)])
    spec = types.SimpleNamespace(role="trainer", monitor_interval=0.0, rdzv_handler=rdzv, max_restarts=2,
                                 get_entrypoint_name=lambda: "train.py")
    agent._worker_group = types.SimpleNamespace(spec=spec, state=WorkerState.HEALTHY, group_rank=0)
    agent._remaining_restarts = 2
    agent._exit_barrier_timeout = 1
    agent._initialize_workers = lambda wg: None

    def monitor(wg):
        st = script[calls["tick"]]
        calls["tick"] += 1
        return RunResult(state=st)

    agent._monitor_workers = monitor
    agent._restart_workers = lambda wg: calls.__setitem__("restart", calls["restart"] + 1)
    agent._stop_workers = lambda wg: None
    agent._exit_barrier = lambda: calls.__setitem__("barrier", calls["barrier"] + 1)
    res = agent._invoke_run()
    assert res.state == WorkerState.SUCCEEDED
    assert calls["restart"] == 2 and calls["barrier"] == 1
    assert agent._remaining_restarts == 1  # only the FAILED tick consumed an attempt
