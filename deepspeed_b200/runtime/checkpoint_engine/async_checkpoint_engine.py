"""Asynchronous tiered checkpoint engine.

Role parity: reference ``nebula_checkpoint_engine.py`` (Azure Nebula async persistence).  Nebula is a
proprietary service, so the equivalent here is self-contained: ``save`` snapshots tensors to host
memory (pinned when CUDA is present) and hands the serialisation + write to a background thread;
``commit`` waits for all writes of the tag and publishes them atomically (rename).  Training resumes
as soon as the device->host snapshot is done.
"""
import os
import queue
import threading

import torch

from deepspeed_b200.utils.logging import logger
from .checkpoint_engine import CheckpointEngine


def _snapshot(obj):
    if torch.is_tensor(obj):
        t = obj.detach()
        if t.device.type != "cpu":
            host = torch.empty(t.shape, dtype=t.dtype, device="cpu", pin_memory=torch.cuda.is_available())
            host.copy_(t, non_blocking=False)
            return host
        return t.clone()
    if isinstance(obj, dict):
        return type(obj)((k, _snapshot(v)) for k, v in obj.items())
    if isinstance(obj, (list, tuple)):
        return type(obj)(_snapshot(v) for v in obj)
    return obj


class AsyncCheckpointEngine(CheckpointEngine):

    def __init__(self, config_params=None, workers=2):
        super().__init__(config_params)
        self._q = queue.Queue()
        self._pending = 0
        self._cv = threading.Condition()
        self._errors = []
        self._threads = [threading.Thread(target=self._run, daemon=True) for _ in range(workers)]
        for t in self._threads:
            t.start()

    def _run(self):
        while True:
            item = self._q.get()
            if item is None:
                return
            sd, path = item
            try:
                tmp = f"{path}.tmp"
                torch.save(sd, tmp)
                os.replace(tmp, path)
            except Exception as e:  # surfaced at commit()
                self._errors.append((path, e))
            finally:
                with self._cv:
                    self._pending -= 1
                    self._cv.notify_all()

    def save(self, state_dict, path: str):
        snap = _snapshot(state_dict)
        with self._cv:
            self._pending += 1
        self._q.put((snap, path))

    def load(self, path: str, map_location=None):
        self.commit(None)
        return torch.load(path, map_location=map_location, weights_only=False)

    def commit(self, tag):
        with self._cv:
            while self._pending > 0:
                self._cv.wait()
        if self._errors:
            errs, self._errors = self._errors, []
            raise RuntimeError(f"async checkpoint writes failed: {errs}")
        if tag is not None:
            logger.debug(f"[Async] Checkpoint {tag} is durable")
        return True
