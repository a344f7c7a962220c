"""Elastic agent: torchelastic ``LocalElasticAgent`` that exports the DeepSpeed env to workers and restarts
the worker group on membership change (reference ``elasticity/elastic_agent.py:32 DSElasticAgent``)."""
import os
from typing import Any, Dict, Optional

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
