"""Experiment scheduler (reference ``autotuning/scheduler.py`` ``ResourceManager``): run each experiment as a
launcher subprocess on a slice of the available GPUs, collect ``metrics.json`` written by the engine."""
import copy
import json
import os
import subprocess
import sys
import time

from deepspeed_b200.utils.logging import logger
from .utils import search_error


class ResourceManager:

    def __init__(self, args, hosts, num_gpus_per_node, results_dir, exps_dir, arg_mappings=None, runner=None):
        self.args = args
        self.hosts, self.num_gpus_per_node = list(hosts), num_gpus_per_node
        self.results_dir, self.exps_dir = results_dir, exps_dir
        self.arg_mappings = arg_mappings or {}
        self.queue, self.finished = [], {}
        self.exp_count = 0
        self._runner = runner  # injectable for tests: fn(exp, result_dir) -> None

    def schedule_experiments_dicts(self, exps):
        for e in exps:
            e = e  # experiments are dicts {name, ds_config, num_gpus, num_nodes}
            e.setdefault("exp_id", self.exp_count)
            self.exp_count += 1
            e["result_dir"] = os.path.join(self.results_dir, e["name"])
            self.queue.append(e)
        return [e["result_dir"] for e in exps]

    def schedule_experiments(self, exp_paths):
        exps = []
        for p in exp_paths:
            with open(p) as f:
                exps.append(json.load(f))
        return self.schedule_experiments_dicts(exps)

    def _cmd(self, exp, cfg_path):
        a = self.args
        user_args = list(a.user_args)
        # point the user script at this experiment's config
        for i, s in enumerate(user_args):
            if s in ("--deepspeed_config", "--deepspeed-config") and i + 1 < len(user_args):
                user_args[i + 1] = cfg_path
            elif s.startswith("--deepspeed_config="):
                user_args[i] = f"--deepspeed_config={cfg_path}"
        for key, arg in self.arg_mappings.items():
            val = exp["ds_config"]
            for part in key.split("."):
                val = val.get(part) if isinstance(val, dict) else None
            if val is not None and arg in user_args:
                user_args[user_args.index(arg) + 1] = str(val)
        return [sys.executable, "-m", "deepspeed_b200.launcher.runner", "--num_gpus", str(exp.get("num_gpus", 1)),
                "--num_nodes", str(exp.get("num_nodes", 1)), "--master_port", str(29600 + exp["exp_id"] % 300),
                a.user_script] + user_args

    def run_job(self, exp):
        rd = exp["result_dir"]
        os.makedirs(rd, exist_ok=True)
        cfg = copy.deepcopy(exp["ds_config"])
        cfg.setdefault("autotuning", {})
        cfg["autotuning"].update({"enabled": True, "metric_path": os.path.join(rd, "metrics.json"),
                                  "model_info_path": os.path.join(rd, "model_info.json")})
        cfg_path = os.path.join(rd, "ds_config.json")
        with open(cfg_path, "w") as f:
            json.dump(cfg, f)
        with open(os.path.join(rd, "exp.json"), "w") as f:
            json.dump({k: v for k, v in exp.items() if k != "result"}, f)
        if self._runner is not None:
            self._runner({**exp, "ds_config": cfg}, rd)
            return
        cmd = self._cmd(exp, cfg_path)
        with open(os.path.join(rd, "stdout.log"), "w") as so, open(os.path.join(rd, "stderr.log"), "w") as se:
            try:
                subprocess.run(cmd, stdout=so, stderr=se, timeout=getattr(self.args, "exp_timeout", 1800))
            except subprocess.TimeoutExpired:
                se.write("Error: experiment timed out\n")

    def run(self):
        while self.queue:
            exp = self.queue.pop(0)
            t = time.time()
            self.run_job(exp)
            err = None
            mp = os.path.join(exp["result_dir"], "metrics.json")
            if not os.path.exists(mp):
                err = search_error(os.path.join(exp["result_dir"], "stderr.log")) or "no metrics produced"
            self.finished[exp["exp_id"]] = (exp, err)
            logger.info(f"exp {exp['name']} done in {time.time() - t:.1f}s" + (f" (error: {err})" if err else ""))

    def metric_of(self, exp, metric):
        mp = os.path.join(exp["result_dir"], "metrics.json")
        if not os.path.exists(mp):
            return None
        with open(mp) as f:
            return json.load(f).get(metric)

    def parse_results(self, metric):
        best, best_val = None, None
        for exp, err in self.finished.values():
            if err:
                continue
            v = self.metric_of(exp, metric)
            if v is None:
                continue
            if best_val is None or (v < best_val if metric == "latency" else v > best_val):
                best, best_val = exp, v
        return best, best_val

    def clear(self):
        self.queue = []
