"""Autotuner: find the ZeRO stage / micro-batch size / ZeRO knobs that maximise throughput (or minimise latency).

Flow parity with reference ``autotuning/autotuner.py`` (``tune :404``): (1) model-info profiling run, (2) for each
ZeRO stage that fits in memory (estimator below), sweep micro-batch sizes, (3) tune the stage's knob space with
the chosen tuner around the best micro-batch size, (4) write the optimal config + a results table.  Memory
arithmetic uses B200's 180 GB HBM via ``torch.cuda.mem_get_info`` when a GPU is visible.
"""
import copy
import json
import os
import shutil

from deepspeed_b200.utils.logging import logger
from . import constants as K
from .config import DeepSpeedAutotuningConfig
from .scheduler import ResourceManager
from .tuner import GridSearchTuner, ModelBasedTuner, RandomTuner
from .utils import canonical_name, get_all_configs, memory_to_string, number_to_string, replace_dict

ZERO_SPACES = {0: K.DEFAULT_TUNING_SPACE_ZERO_0, 1: K.DEFAULT_TUNING_SPACE_ZERO_1, 2: K.DEFAULT_TUNING_SPACE_ZERO_2,
               3: K.DEFAULT_TUNING_SPACE_ZERO_3}


class Autotuner:

    def __init__(self, args, active_resources, runner=None):
        self.args = args
        self.user_config = self._get_user_config(args.user_args)
        assert self.user_config is not None, "DeepSpeed configuration is not provided"
        self.autotuning_config = DeepSpeedAutotuningConfig(**(self.user_config.get(K.AUTOTUNING) or {}))
        self.results_dir, self.exps_dir = self.autotuning_config.results_dir, self.autotuning_config.exps_dir
        if self.autotuning_config.overwrite:
            for d in (self.results_dir, self.exps_dir):
                shutil.rmtree(d, ignore_errors=True)
        os.makedirs(self.results_dir, exist_ok=True)
        os.makedirs(self.exps_dir, exist_ok=True)
        self.exp_num_nodes = len(active_resources)
        self.exp_num_gpus = min(len(v) for v in active_resources.values())
        self.rm = ResourceManager(args, list(active_resources.keys()), self.exp_num_gpus, self.results_dir, self.exps_dir,
                                  self.autotuning_config.arg_mappings, runner=runner)
        self.records = {}
        self.optimal_cmd = None
        self.optimal_ds_config = None
        self.model_info = self.autotuning_config.model_info

    # ---- config plumbing
    def _get_user_config(self, user_args):
        path = None
        for i, a in enumerate(user_args):
            if a in ("--deepspeed_config", "--deepspeed-config") and i + 1 < len(user_args):
                path = user_args[i + 1]
            elif a.startswith("--deepspeed_config="):
                path = a.split("=", 1)[1]
        if path is None:
            return None
        assert os.path.isfile(path), f"DeepSpeed configuration file: {path} is not an existing file"
        self.user_config_path = path
        with open(path) as f:
            return json.load(f)

    def metric(self):
        return self.autotuning_config.metric

    def fast_enabled(self):
        return self.autotuning_config.fast

    def mp_size(self):
        return self.autotuning_config.mp_size

    def get_gpu_memory_info(self):
        try:
            import torch
            if torch.cuda.is_available():
                return torch.cuda.get_device_properties(0).total_memory
        except Exception:
            pass
        return 180 * (1 << 30)

    def get_model_num_params(self):
        return (self.model_info or {}).get(K.MODEL_INFO_NUM_PARAMS)

    def get_instantiation_memory_required_per_gpu(self, zero_stage):
        """params(2B) + grads(2B) + optimizer(12B: fp32 master + 2 moments), divided per the stage's sharding."""
        n = self.get_model_num_params() or 0
        g = self.exp_num_gpus * self.exp_num_nodes // self.mp_size()
        p, gr, o = 2 * n, 2 * n, 12 * n
        if zero_stage >= 1:
            o /= g
        if zero_stage >= 2:
            gr /= g
        if zero_stage >= 3:
            p /= g
        return (p + gr + o) / self.mp_size()

    # ---- phases
    def model_info_profile_run(self):
        if self.model_info and self.model_info.get(K.MODEL_INFO_NUM_PARAMS):
            return self.model_info
        cfg = copy.deepcopy(self.user_config)
        replace_dict(cfg, K.DEFAULT_MIN_MEM_CONFIG)
        cfg.setdefault(K.AUTOTUNING, {})["model_info"] = {K.MODEL_INFO_PROFILE: True}
        exp = {"name": "profile_model_info", "ds_config": cfg, "num_gpus": self.exp_num_gpus, "num_nodes": self.exp_num_nodes}
        self.rm.schedule_experiments_dicts([exp])
        self.rm.run()
        p = os.path.join(exp["result_dir"], "model_info.json")
        if os.path.exists(p):
            with open(p) as f:
                self.model_info = json.load(f)
        self.rm.clear()
        return self.model_info

    def _mbs_candidates(self, stage):
        c = self.autotuning_config
        lo, hi = c.min_train_micro_batch_size_per_gpu, c.max_train_micro_batch_size_per_gpu
        user = self.user_config.get("train_micro_batch_size_per_gpu")
        if isinstance(user, int):
            return [user]
        out, m = [], max(1, lo)
        while m <= hi and len(out) < max(1, c.num_tuning_micro_batch_sizes):
            out.append(m)
            m *= 2
        return out

    def _exp(self, space_name, ds_config):
        name = canonical_name(ds_config, ["stage", "train_micro_batch_size_per_gpu", "overlap_comm", "b200_unit_prefetch",
                                          "reduce_bucket_size", "stage3_param_persistence_threshold"], prefix=space_name)
        return {"name": name, "ds_config": ds_config, "num_gpus": self.exp_num_gpus, "num_nodes": self.exp_num_nodes}

    def tune_space(self, stage):
        space_name = f"{K.TUNING_MICRO_BATCH_SIZE_PREFIX}{stage}"
        base = copy.deepcopy(self.user_config)
        base.pop(K.AUTOTUNING, None)
        # 1) micro-batch sweep with the first value of every knob
        from .utils import get_first_config
        first = get_first_config(ZERO_SPACES[stage])
        best_mbs, best_val = None, None
        for mbs in self._mbs_candidates(stage):
            cfg = replace_dict(copy.deepcopy(base), first)
            cfg["train_micro_batch_size_per_gpu"] = mbs
            cfg.pop("train_batch_size", None)
            exp = self._exp(space_name, cfg)
            self.rm.schedule_experiments_dicts([exp])
            self.rm.run()
            val = self.rm.metric_of(exp, self.metric())
            self.update_records(space_name, exp, val, 1)
            self.rm.clear()
            if val is None:
                break  # larger batches will not fit either
            if best_val is None or (val < best_val if self.metric() == "latency" else val > best_val):
                best_mbs, best_val = mbs, val
            elif self.fast_enabled():
                break  # throughput plateaued
        if best_mbs is None:
            return None
        if self.fast_enabled():
            return self.get_best_space_record(space_name)
        # 2) knob space around the best micro batch
        exps = []
        for cfg_knobs in get_all_configs(ZERO_SPACES[stage]):
            cfg = replace_dict(copy.deepcopy(base), cfg_knobs)
            cfg["train_micro_batch_size_per_gpu"] = best_mbs
            cfg.pop("train_batch_size", None)
            exps.append(self._exp(space_name, cfg))
        t = self.autotuning_config.tuner_type
        tuner_cls = {K.AUTOTUNING_TUNER_GRIDSEARCH: GridSearchTuner, K.AUTOTUNING_TUNER_RANDOM: RandomTuner,
                     K.AUTOTUNING_TUNER_MODELBASED: ModelBasedTuner}[t]
        tuner = tuner_cls(exps, self.rm, self.metric())
        n = tuner.tune(sample_size=1, n_trials=self.autotuning_config.tuner_num_trials,
                       early_stopping=self.autotuning_config.tuner_early_stopping)
        if tuner.best_exp is not None:
            self.update_records(space_name, tuner.best_exp, tuner.best_metric_val, n)
        return self.get_best_space_record(space_name)

    def tune(self):
        self.model_info_profile_run()
        gpu_mem = self.get_gpu_memory_info()
        n = self.get_model_num_params()
        if n:
            logger.info(f"The model has {number_to_string(n)} parameters; device memory {memory_to_string(gpu_mem, 'B')}")
        stages = self.autotuning_config.zero_stages
        user_stage = self.user_config.get("zero_optimization", {}).get("stage")
        if stages is None:
            stages = [user_stage] if isinstance(user_stage, int) else [0, 1, 2, 3]
        for s in stages:
            need = self.get_instantiation_memory_required_per_gpu(s) if n else 0
            if need > gpu_mem:
                logger.info(f"ZeRO stage {s}: needs {memory_to_string(need, 'B')} per GPU for model states alone, skipping")
                continue
            self.tune_space(s)
        best = self.get_best_space_records()
        if best.get(K.GLOBAL_TUNING_SPACE):
            exp, val, _ = best[K.GLOBAL_TUNING_SPACE]
            self.optimal_ds_config = exp["ds_config"]
        return best

    # ---- records / reporting
    def update_records(self, space_name, exp, metric_val, num_exps):
        self.records.setdefault(space_name, []).append((exp, metric_val, num_exps))

    def get_best_space_record(self, space_name):
        recs = [r for r in self.records.get(space_name, []) if r[1] is not None]
        if not recs:
            return None
        key = (lambda r: -r[1]) if self.metric() != "latency" else (lambda r: r[1])
        best = sorted(recs, key=key)[0]
        return (best[0], best[1], sum(r[2] for r in self.records[space_name]))

    def get_best_space_records(self):
        out, glob = {}, None
        for name in self.records:
            b = self.get_best_space_record(name)
            if b is None:
                continue
            out[name] = b
            if glob is None or (b[1] < glob[1] if self.metric() == "latency" else b[1] > glob[1]):
                glob = b
        if glob is not None:
            out[K.GLOBAL_TUNING_SPACE] = glob
        return out

    def print_tuning_results(self):
        best = self.get_best_space_records()
        rows = [(name, n, val, exp["name"]) for name, (exp, val, n) in best.items()]
        try:
            from tabulate import tabulate
            print(tabulate(rows, headers=["tuning_space", "num_experiments", "best_metric_val", "best_exp_name"],
                           tablefmt="pipe"))
        except Exception:
            for r in rows:
                print(r)
        if K.GLOBAL_TUNING_SPACE in best:
            exp, val, n = best[K.GLOBAL_TUNING_SPACE]
            print(f"\nTuning completed. Best {self.metric()} = {val} from {exp['name']} after {n} experiments.")

    def write_optimal_config(self):
        best = self.get_best_space_records().get(K.GLOBAL_TUNING_SPACE)
        if not best:
            return
        exp, _, _ = best
        cfg = copy.deepcopy(exp["ds_config"])
        cfg.pop(K.AUTOTUNING, None)
        path = os.path.join(self.results_dir, "ds_config_optimal.json")
        with open(path, "w") as f:
            json.dump(cfg, f, indent=2)
        user_args = list(self.args.user_args)
        for i, a in enumerate(user_args):
            if a in ("--deepspeed_config", "--deepspeed-config") and i + 1 < len(user_args):
                user_args[i + 1] = path
        self.optimal_cmd = ["deepspeed", self.args.user_script] + user_args
        with open(os.path.join(self.results_dir, "cmd_optimal.txt"), "w") as f:
            f.write(" ".join(self.optimal_cmd))
        self.optimal_ds_config = cfg
        logger.info(f"Wrote optimal config to {path}")

    def run_after_tuning(self):
        if self.optimal_cmd:
            import subprocess
            import sys
            cmd = [sys.executable, "-m", "deepspeed_b200.launcher.runner"] + self.optimal_cmd[1:]
            subprocess.Popen(cmd).wait()
        else:
            logger.info("No optimal DeepSpeed configuration found by autotuning.")
