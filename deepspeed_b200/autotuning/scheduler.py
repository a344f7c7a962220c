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


class Node:
    """GPU slots of one host (reference ``scheduler.py:259``)."""

    def __init__(self, host, max_slots):
        self.host, self.max_slots = host, max_slots
        self.idle_slots = list(range(max_slots))

    def reserve_slots(self, slot_request: int):
        if len(self.idle_slots) < slot_request:
            return None
        taken, self.idle_slots = self.idle_slots[:slot_request], self.idle_slots[slot_request:]
        return taken

    def restore_slots(self, slots):
        self.idle_slots = sorted(set(self.idle_slots) | set(slots))


class Reservation:
    """Slots held on one node for the lifetime of one experiment."""

    def __init__(self, node, slots):
        self.node, self.slots = node, slots

    def restore_slots(self):
        self.node.restore_slots(self.slots)

    def desc(self):
        return f"{self.node.host}:{','.join(map(str, sorted(self.slots)))}@"


def get_job_id():
    """Cluster job id when the scheduler exports one."""
    for var in ("DLWS_JOB_ID", "DLTS_JOB_ID", "SLURM_JOB_ID"):
        if var in os.environ:
            return os.environ[var]
    return "unknown-job-id"


def get_user():
    return os.environ.get("USER", "unknown-user")


def include_string(reservations):
    """``host:0,1@host2:0,1`` for the launcher's ``--include``."""
    return "@".join(r.desc().rstrip("@") for r in reservations)


class ResourceManager:

    def __init__(self, args, hosts, num_gpus_per_node, results_dir, exps_dir, arg_mappings=None, runner=None):
        self.args = args
        self.hosts, self.num_gpus_per_node = list(hosts), num_gpus_per_node
        self.nodes = [Node(h, num_gpus_per_node) for h in self.hosts]
        self.running = {}  # exp_id -> (thread, exp, reservations, t0)
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
        where = ["--num_gpus", str(exp.get("num_gpus", 1)), "--num_nodes", str(exp.get("num_nodes", 1))]
        res = exp.get("reservations")
        if res and (len(self.nodes) > 1 or res[0].node.host not in ("localhost", "127.0.0.1")):
            where = ["--include", include_string(res)]
            if getattr(a, "hostfile", None):
                where = ["--hostfile", a.hostfile] + where
        elif res:
            exp["visible_devices"] = ",".join(map(str, sorted(res[0].slots)))
        return [sys.executable, "-m", "deepspeed_b200.launcher.runner"] + where + [
            "--master_port", str(29600 + exp["exp_id"] % 300), a.user_script] + user_args

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
            json.dump({k: v for k, v in exp.items() if k not in ("result", "reservations")}, f)
        if self._runner is not None:
            self._runner({**exp, "ds_config": cfg}, rd)
            return
        cmd = self._cmd(exp, cfg_path)
        with open(os.path.join(rd, "stdout.log"), "w") as so, open(os.path.join(rd, "stderr.log"), "w") as se:
            env = os.environ.copy()
            if exp.get("visible_devices") is not None:
                env["CUDA_VISIBLE_DEVICES"] = exp["visible_devices"]
            try:
                subprocess.run(cmd, stdout=so, stderr=se, env=env, timeout=getattr(self.args, "exp_timeout", 1800))
            except subprocess.TimeoutExpired:
                se.write("Error: experiment timed out\n")

    # ---- resources ------------------------------------------------------------------------------------------------------
    def resource_request(self, exp):
        return exp.get("num_nodes", 1), exp.get("num_gpus", 1)

    def request_resource(self, exp):
        """Reserve ``num_gpus`` slots on ``num_nodes`` nodes, or nothing at all."""
        n_nodes, n_slots = self.resource_request(exp)
        got = []
        for node in self.nodes:
            if len(got) == n_nodes:
                break
            slots = node.reserve_slots(n_slots)
            if slots is not None:
                got.append(Reservation(node, slots))
        if len(got) < n_nodes:
            for r in got:
                r.restore_slots()
            return None
        return got

    def status(self):
        return " ".join(f"{n.host} ({len(n.idle_slots)} idle gpus)" for n in self.nodes)

    def _finish(self, exp, reservations, t0):
        for r in reservations:
            r.restore_slots()
        err = None
        mp = os.path.join(exp["result_dir"], "metrics.json")
        if not os.path.exists(mp):
            err = search_error(os.path.join(exp["result_dir"], "stderr.log")) or "no metrics produced"
        self.finished[exp["exp_id"]] = (exp, err)
        logger.info(f"exp {exp['name']} done in {time.time() - t0:.1f}s" + (f" (error: {err})" if err else ""))

    def _reap(self):
        for eid, (th, exp, res, t0) in list(self.running.items()):
            if not th.is_alive():
                th.join()
                del self.running[eid]
                self._finish(exp, res, t0)

    def run(self):
        """Start every queued experiment as soon as enough slots are free; experiments that fit side by side run
        concurrently (reference ``ResourceManager.run``)."""
        import threading
        max_nodes, max_slots = len(self.nodes), self.num_gpus_per_node
        while self.queue or self.running:
            self._reap()
            started = False
            if self.queue:
                exp = self.queue[0]
                n_nodes, n_slots = self.resource_request(exp)
                if n_nodes > max_nodes or n_slots > max_slots:
                    self.queue.pop(0)
                    self.finished[exp["exp_id"]] = (exp, f"needs {n_nodes}x{n_slots} GPUs, pool is {max_nodes}x{max_slots}")
                    continue
                res = self.request_resource(exp)
                if res is not None:
                    self.queue.pop(0)
                    exp["reservations"] = res
                    th = threading.Thread(target=self.run_job, args=(exp, ), daemon=True)
                    self.running[exp["exp_id"]] = (th, exp, res, time.time())
                    th.start()
                    started = True
            if not started:
                time.sleep(0.01 if self._runner is not None else 0.5)

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
        """Drop queued work and wait for what is running."""
        self.queue = []
        while self.running:
            self._reap()
            time.sleep(0.01)


def run_experiment(exp: dict, reservations, user_script, user_args):
    """Run ONE experiment on the given reservations with a throw-away manager (reference ``scheduler.py:310``)."""
    import types
    hosts = [r.node.host for r in reservations]
    args = types.SimpleNamespace(user_script=user_script, user_args=list(user_args), hostfile=exp.get("hostfile"))
    rm = ResourceManager(args, hosts, max(len(r.slots) for r in reservations), os.path.dirname(exp["result_dir"]), None)
    exp = dict(exp, reservations=reservations)
    exp.setdefault("exp_id", 0)
    rm.run_job(exp)
    return exp


def clean_up(exp: dict, reservations):
    """Stop what an experiment left behind.  Experiments are launcher subprocesses that own a process group and the
    launcher forwards termination to every node it started, so signalling that exact group (recorded in ``exp["pgid"]``
    when the experiment was started detached) is sufficient -- nothing is ever matched by command-line pattern."""
    import signal
    pgid = exp.get("pgid")
    if pgid:
        try:
            os.killpg(int(pgid), signal.SIGTERM)
        except (ProcessLookupError, PermissionError):
            pass
    for r in reservations:
        r.restore_slots()
