"""Synchronous ``torch.save`` / ``torch.load`` engine (reference: ``torch_checkpoint_engine.py``)."""
import torch

from deepspeed_b200.utils.logging import logger
from .checkpoint_engine import CheckpointEngine


class TorchCheckpointEngine(CheckpointEngine):

    def create(self, tag):
        logger.debug(f"[Torch] Checkpoint {tag} is about to be saved!")

    def save(self, state_dict, path: str):
        tmp = f"{path}.tmp"
        torch.save(state_dict, tmp)
        import os
        os.replace(tmp, path)  # atomic publish: a crash never leaves a truncated shard

    def load(self, path: str, map_location=None):
        try:
            return torch.load(path, map_location=map_location, weights_only=False)
        except ModuleNotFoundError as e:
            if not str(e.name or "").startswith("deepspeed"):
                raise
            # a checkpoint written by upstream DeepSpeed pickles a few of its own classes (LossScaler, ZeroStageEnum, ...):
            # resolve ``deepspeed.*`` to the same-path module of this package
            return torch.load(path, map_location=map_location, weights_only=False, pickle_module=_compat_pickle())

    def commit(self, tag):
        logger.debug(f"[Torch] Checkpoint {tag} is ready now!")
        return True


def _compat_pickle():
    """A ``pickle``-compatible module whose Unpickler maps ``deepspeed.<x>`` to ``deepspeed_b200.<x>`` (same module paths
    and class names are kept on purpose); unknown attributes degrade to an inert placeholder."""
    import importlib
    import pickle
    import types

    class _Placeholder:

        def __init__(self, *a, **k):
            pass

        def __setstate__(self, state):
            if isinstance(state, dict):
                self.__dict__.update(state)

    class _Unpickler(pickle.Unpickler):

        def find_class(self, module, name):
            if module == "deepspeed" or module.startswith("deepspeed."):
                alt = "deepspeed_b200" + module[len("deepspeed"):]
                try:
                    return getattr(importlib.import_module(alt), name)
                except (ImportError, AttributeError):
                    return type(name, (_Placeholder, ), {})
            return super().find_class(module, name)

    mod = types.ModuleType("dsb200_compat_pickle")
    mod.Unpickler = _Unpickler
    mod.load = lambda f, **kw: _Unpickler(f, **kw).load()
    mod.loads = pickle.loads
    mod.dump, mod.dumps, mod.Pickler = pickle.dump, pickle.dumps, pickle.Pickler
    mod.__name__ = "pickle"
    return mod
