"""Tiered checkpoint engine (role of reference ``runtime/checkpoint_engine/nebula_checkpoint_engine.py``).

The reference delegates to the proprietary ``torch_nebula`` service: checkpoints are snapshotted to fast local storage,
persisted to slow durable storage in the background at most every ``persistent_time_interval`` seconds, and only the
newest ``num_of_version_in_retention`` versions are kept.  This engine implements those semantics itself:

* ``save``    - host snapshot + asynchronous write to the (fast) path the trainer asked for (``AsyncCheckpointEngine``);
* ``commit``  - waits for the tag's files, then mirrors the tag directory to ``persistent_storage_path`` on a
  background thread if the persistence interval elapsed, and prunes old versions there;
* ``load``    - prefers the fast tier and falls back to the persistent tier (``enable_nebula_load``).
"""
import os
import shutil
import threading
import time

from deepspeed_b200.utils import logger

from .async_checkpoint_engine import AsyncCheckpointEngine


class NebulaCheckpointEngine(AsyncCheckpointEngine):

    def __init__(self, config_params=None, workers=2):
        super().__init__(config_params, workers=workers)
        nc = getattr(config_params, "nebula_config", config_params)
        self.persist_root = getattr(nc, "persistent_storage_path", None)
        self.persist_interval = float(getattr(nc, "persistent_time_interval", 100))
        self.retention = int(getattr(nc, "num_of_version_in_retention", 2))
        self.enable_load = bool(getattr(nc, "enable_nebula_load", True))
        self.load_path = getattr(nc, "load_path", None)
        self._tag_dirs = {}  # tag -> fast-tier directory holding its files
        self._tag = None
        self._last_persist = 0.0
        self._mirror = None
        if self.persist_root:
            os.makedirs(self.persist_root, exist_ok=True)

    def create(self, tag):
        self._tag = str(tag)
        logger.info(f"[Nebula] start checkpoint version {tag}")

    def save(self, state_dict, path: str):
        if self._tag is not None:
            self._tag_dirs.setdefault(self._tag, os.path.dirname(path))
        super().save(state_dict, path)

    def _persisted_versions(self):
        if not self.persist_root:
            return []
        vs = [d for d in os.listdir(self.persist_root) if os.path.isdir(os.path.join(self.persist_root, d))]
        return sorted(vs, key=lambda d: os.path.getmtime(os.path.join(self.persist_root, d)))

    def _mirror_tag(self, tag, src):
        dst = os.path.join(self.persist_root, tag)
        tmp = dst + ".partial"
        try:
            shutil.rmtree(tmp, ignore_errors=True)
            shutil.copytree(src, tmp)
            shutil.rmtree(dst, ignore_errors=True)
            os.replace(tmp, dst)
            with open(os.path.join(self.persist_root, "latest"), "w") as f:
                f.write(tag)
            for old in self._persisted_versions()[:-self.retention] if self.retention > 0 else []:
                shutil.rmtree(os.path.join(self.persist_root, old), ignore_errors=True)
            logger.info(f"[Nebula] version {tag} persisted to {dst}")
        except Exception as e:  # persistence is best effort; the fast tier still has the checkpoint
            logger.warning(f"[Nebula] persisting {tag} failed: {e}")

    def commit(self, tag):
        super().commit(tag)
        tag = str(tag) if tag is not None else self._tag
        src = self._tag_dirs.pop(tag, None) if tag is not None else None
        now = time.time()
        import torch.distributed as td
        rank0 = True
        if td.is_available() and td.is_initialized():
            td.barrier()  # every rank's files of this version are on disk before rank 0 mirrors the directory
            rank0 = td.get_rank() == 0
        if rank0 and self.persist_root and src and os.path.isdir(src) and (now - self._last_persist >= self.persist_interval
                                                                 or self._last_persist == 0.0):
            self.wait_persisted()
            self._last_persist = now
            self._mirror = threading.Thread(target=self._mirror_tag, args=(tag, src), daemon=True)
            self._mirror.start()
        return True

    def wait_persisted(self):
        if self._mirror is not None:
            self._mirror.join()
            self._mirror = None

    def load(self, path: str, map_location=None):
        if os.path.isfile(path) or not self.enable_load:
            return super().load(path, map_location)
        self.wait_persisted()
        root = self.load_path or self.persist_root
        if root:
            tag_dir, fname = os.path.split(path)
            cand = os.path.join(root, os.path.basename(tag_dir), fname)
            if os.path.isfile(cand):
                logger.info(f"[Nebula] {path} not in the fast tier; loading {cand}")
                return super().load(cand, map_location)
        return super().load(path, map_location)
