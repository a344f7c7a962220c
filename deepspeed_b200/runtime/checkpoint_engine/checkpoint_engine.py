"""Checkpoint engine interface (reference: ``checkpoint_engine.py:9``): create/save/load/commit/makedirs."""
import os


class CheckpointEngine:

    def __init__(self, config_params=None):
        self.config = config_params

    def create(self, tag):
        pass

    def makedirs(self, path, exist_ok=False):
        os.makedirs(path, exist_ok=exist_ok)

    def save(self, state_dict, path: str):
        raise NotImplementedError

    def load(self, path: str, map_location=None):
        raise NotImplementedError

    def commit(self, tag):
        return True
