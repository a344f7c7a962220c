"""Elastic agent: torchelastic ``LocalElasticAgent`` that exports the DeepSpeed env to workers and supervises the worker
group -- restart on failure (bounded by ``max_restarts``), on lost heartbeats of peer nodes and on membership growth
(new nodes waiting at the rendezvous; these do not consume restart attempts).  Reference: ``elasticity/elastic_agent.py:32
DSElasticAgent`` and its run loop ``:127 _invoke_run``.

The policy is a pure function (:func:`supervise_decision`) so it is unit-tested without torchelastic; the loop only gathers
the observations (worker-group state, rendezvous participants, nodes waiting, heartbeats) and applies the decision.
"""
import os
import time
from datetime import datetime, timedelta
from typing import Any, Dict, Optional

FINISH, RESTART, RESCALE, FAIL, CONTINUE = "finish", "restart", "rescale", "fail", "continue"


def supervise_decision(state: str, remaining_restarts: int, participants_before: int, participants_now: int,
                       nodes_waiting: int, dead_nodes: int = 0) -> str:
    """What the agent does after one monitoring tick.

    ``state``: worker-group state name (SUCCEEDED / FAILED / UNHEALTHY / HEALTHY).  A shrunken rendezvous or nodes whose
    heartbeat expired are treated like a failure (the collective job cannot continue); new nodes waiting while healthy
    trigger a re-rendezvous that is NOT charged to the restart budget."""
    state = state.upper()
    if state == "SUCCEEDED":
        return FINISH
    lost = participants_now < participants_before or dead_nodes > 0
    if state in ("FAILED", "UNHEALTHY") or lost:
        return RESTART if remaining_restarts > 0 else FAIL
    if state == "HEALTHY":
        return RESCALE if nodes_waiting > 0 else CONTINUE
    raise RuntimeError(f"worker group in unexpected state {state}")

try:
    from torch.distributed.elastic.agent.server.local_elastic_agent import LocalElasticAgent
    from torch.distributed.elastic.agent.server.api import WorkerSpec
    _HAVE_ELASTIC = True
except Exception:  # pragma: no cover
    LocalElasticAgent = object
    WorkerSpec = object
    _HAVE_ELASTIC = False


class DSElasticAgent(LocalElasticAgent):

    def __init__(self, spec, env: Dict, start_method="spawn", exit_barrier_timeout: float = 300,
                 log_dir: Optional[str] = None, logs_specs=None):
        if not _HAVE_ELASTIC:
            raise RuntimeError("torch.distributed.elastic is unavailable")
        kwargs = dict(start_method=start_method, exit_barrier_timeout=exit_barrier_timeout)
        if logs_specs is not None:
            kwargs["logs_specs"] = logs_specs
        else:
            try:
                from torch.distributed.elastic.multiprocessing import DefaultLogsSpecs
                kwargs["logs_specs"] = DefaultLogsSpecs(log_dir=log_dir)
            except Exception:
                kwargs["log_dir"] = log_dir
        super().__init__(spec, **kwargs)
        self.ds_env = dict(env)

    @staticmethod
    def _set_master_addr_port(store, master_addr: Optional[str], master_port: Optional[int], local_addr=None):
        import socket
        if master_port is None:
            with socket.socket() as s:
                s.bind(("", 0))
                master_port = s.getsockname()[1]
        if master_addr is None:
            master_addr = local_addr or "127.0.0.1"
        store.set("MASTER_ADDR", master_addr.encode("utf-8"))
        store.set("MASTER_PORT", str(master_port).encode("utf-8"))

    def _start_workers(self, worker_group) -> Dict[int, Any]:
        # make the DeepSpeed launcher env (NCCL knobs, elasticity config, PYTHONPATH...) visible to every worker
        for k, v in self.ds_env.items():
            os.environ.setdefault(k, str(v))
        return super()._start_workers(worker_group)

    # ---- observations ------------------------------------------------------------------------------------------------
    @staticmethod
    def _rdzv_state(rdzv_handler):
        holder = getattr(rdzv_handler, "_state_holder", None)
        return getattr(holder, "state", None)

    def _participants(self, rdzv_handler) -> int:
        st = self._rdzv_state(rdzv_handler)
        return len(getattr(st, "participants", {}) or {})

    def _dead_nodes(self, rdzv_handler) -> int:
        st, settings = self._rdzv_state(rdzv_handler), getattr(rdzv_handler, "_settings", None)
        beats = getattr(st, "last_heartbeats", None)
        if not beats or settings is None:
            return 0
        ttl = settings.keep_alive_interval * settings.keep_alive_max_attempt
        if not isinstance(ttl, timedelta):
            ttl = timedelta(seconds=float(ttl))
        sample = next(iter(beats.values()))
        now = datetime.now(sample.tzinfo) if getattr(sample, "tzinfo", None) else datetime.utcnow()
        return sum(1 for t in beats.values() if t < now - ttl)

    # ---- run loop (reference elastic_agent.py:127) ----------------------------------------------------------------------
    def _invoke_run(self, role: str = "default"):
        from torch.distributed.elastic.agent.server.api import WorkerState
        from torch.distributed.elastic.metrics import put_metric
        from deepspeed_b200.utils.logging import logger
        spec = self._worker_group.spec
        role = spec.role
        logger.info(f"[{role}] starting workers for entrypoint: {spec.get_entrypoint_name()}")
        self._initialize_workers(self._worker_group)
        rdzv = spec.rdzv_handler
        members = self._participants(rdzv)
        while True:
            assert self._worker_group.state != WorkerState.INIT
            time.sleep(spec.monitor_interval)
            result = self._monitor_workers(self._worker_group)
            self._worker_group.state = result.state
            put_metric(f"workers.{role}.remaining_restarts", self._remaining_restarts)
            put_metric(f"workers.{role}.{result.state.name.lower()}", 1)
            try:
                waiting = rdzv.num_nodes_waiting()
            except Exception:
                waiting = 0
            action = supervise_decision(result.state.name, self._remaining_restarts, members, self._participants(rdzv),
                                        waiting, self._dead_nodes(rdzv))
            if action == FINISH:
                logger.info(f"[{role}] worker group finished; waiting {self._exit_barrier_timeout}s for the other agents")
                self._exit_barrier()
                return result
            if action == RESTART:
                logger.info(f"[{role}] worker group {result.state.name}: {self._remaining_restarts}/{spec.max_restarts} "
                            f"restarts left, restarting")
                self._remaining_restarts -= 1
                self._restart_workers(self._worker_group)
                members = self._participants(rdzv)
            elif action == RESCALE:
                logger.info(f"[{role}] {waiting} new node(s) at the rendezvous (group_rank={self._worker_group.group_rank}): "
                            f"re-forming the worker group")
                self._restart_workers(self._worker_group)  # membership changes are not charged to max_restarts
                members = self._participants(rdzv)
            elif action == FAIL:
                self._stop_workers(self._worker_group)
                self._worker_group.state = WorkerState.FAILED
                self._exit_barrier()
                return result
