import json
import os
from types import SimpleNamespace

import pytest

from deepspeed_b200.autotuning import Autotuner
from deepspeed_b200.autotuning.tuner import GridSearchTuner, ModelBasedTuner, RandomTuner, RidgeCostModel
from deepspeed_b200.autotuning.utils import canonical_name, get_all_configs, replace_dict


def test_space_expansion():
    space = {"zero_optimization": {"stage": 2, "overlap_comm": [True, False], "reduce_bucket_size": [1, 2, 3]}}
    all_cfg = get_all_configs(space)
    assert len(all_cfg) == 6 and all(c["zero_optimization"]["stage"] == 2 for c in all_cfg)
    assert replace_dict({"a": {"b": 1, "c": 2}}, {"a": {"b": None, "d": 3}}) == {"a": {"c": 2, "d": 3}}
    assert canonical_name({"zero_optimization": {"stage": 2}, "train_micro_batch_size_per_gpu": 4},
                          ["stage", "train_micro_batch_size_per_gpu"], prefix="z2") == "z2_tmbspg4_zos2"


def _fake_runner(perf):
    """Simulated cluster: throughput model + OOM above a micro-batch limit depending on the stage."""

    def run(exp, rd):
        cfg = exp["ds_config"]
        at = cfg["autotuning"]
        if (at.get("model_info") or {}).get("profile"):
            with open(at["model_info_path"], "w") as f:
                json.dump({"num_params": 1_000_000_000}, f)
            return
        z = cfg.get("zero_optimization", {})
        stage, mbs = z.get("stage", 0), cfg["train_micro_batch_size_per_gpu"]
        if mbs > {0: 2, 1: 4, 2: 8, 3: 16}[stage]:
            with open(os.path.join(rd, "stderr.log"), "w") as f:
                f.write("RuntimeError: CUDA out of memory\n")
            return
        tput = perf(stage, mbs, z)
        with open(at["metric_path"], "w") as f:
            json.dump({"throughput": tput, "latency": 1000.0 * mbs / tput}, f)

    return run


@pytest.mark.parametrize("tuner", ["gridsearch", "random", "model_based"])
def test_autotuner_end_to_end(tmp_path, tuner):
    cfg = {"train_micro_batch_size_per_gpu": "auto", "optimizer": {"type": "AdamW"},
           "autotuning": {"enabled": True, "fast": False, "tuner_type": tuner, "results_dir": str(tmp_path / "res"),
                          "exps_dir": str(tmp_path / "exps"), "num_tuning_micro_batch_sizes": 6, "tuner_num_trials": 40,
                          "tuner_early_stopping": 40}}
    path = tmp_path / "ds.json"
    path.write_text(json.dumps(cfg))
    args = SimpleNamespace(user_script="train.py", user_args=["--deepspeed_config", str(path)])

    def perf(stage, mbs, z):
        base = 100.0 * mbs / (1 + 0.05 * mbs) * (1 - 0.05 * stage)
        if stage == 3:
            base *= 1.0 + 0.02 * z.get("b200_unit_prefetch", 1) + (0.05 if z.get("overlap_comm", True) else 0.0)
        return base

    at = Autotuner(args, {"localhost": [0, 1]}, runner=_fake_runner(perf))
    best = at.tune()
    at.print_tuning_results()
    at.write_optimal_config()
    g_exp, g_val, _ = best["global"]
    z = g_exp["ds_config"]["zero_optimization"]
    assert z["stage"] == 3 and g_exp["ds_config"]["train_micro_batch_size_per_gpu"] == 16
    if tuner == "gridsearch":
        assert z["b200_unit_prefetch"] == 4 and z["overlap_comm"] is True
    opt = json.loads((tmp_path / "res" / "ds_config_optimal.json").read_text())
    assert "autotuning" not in opt and opt["zero_optimization"]["stage"] == 3
    assert at.get_instantiation_memory_required_per_gpu(3) < at.get_instantiation_memory_required_per_gpu(0)


def test_cost_model_ranks():
    import numpy as np
    rng = np.random.default_rng(0)
    X = rng.uniform(0, 1, (40, 3))
    y = 3 * X[:, 0] - 2 * X[:, 1] * X[:, 2] + 0.5
    m = RidgeCostModel()
    m.fit(X[:30], y[:30])
    pred = m.predict(X[30:])
    assert np.corrcoef(pred, y[30:])[0, 1] > 0.98


@pytest.mark.parametrize("loss", ["reg", "rank"])
def test_boosted_trees_cost_model_ranks_nonlinear_surface(loss):
    """The in-tree stand-in for the reference's XGBoost cost model: depth-3 boosted trees must rank a throughput surface
    with a threshold (OOM-like cliff) that a linear model cannot express."""
    import numpy as np
    from deepspeed_b200.autotuning.tuner import BoostedTreesCostModel
    rng = np.random.default_rng(1)
    X = rng.uniform(0, 1, (80, 4))
    y = 100 * X[:, 0] * (X[:, 1] < 0.6) + 20 * X[:, 2] + 5.0
    m = BoostedTreesCostModel(loss)
    m.fit(X[:60], y[:60])
    pred = m.predict(X[60:])
    order_true, order_pred = np.argsort(np.argsort(y[60:])), np.argsort(np.argsort(pred))
    rho = np.corrcoef(order_true, order_pred)[0, 1]
    assert rho > 0.8, rho
    assert int(np.argmax(pred)) in set(np.argsort(y[60:])[-3:].tolist())
