"""Persistence back-end protocol (reference ``runtime/checkpoint_engine/checkpoint_engine.py:9``).

Life cycle of one checkpoint version: ``create(tag)`` -> any number of ``save(state, path)`` -> ``commit(tag)`` (durable
when it returns True).  ``load`` and ``makedirs`` are stateless helpers so engines can redirect storage.
"""
import abc
import os


class CheckpointEngine(abc.ABC):

    def __init__(self, config_params=None):
        self.config = config_params

    # -- version life cycle -----------------------------------------------------------------------------------------
    def create(self, tag):
        """Start version ``tag`` (no-op for stateless engines)."""

    @abc.abstractmethod
    def save(self, state_dict, path: str):
        ...

    def commit(self, tag):
        """Everything saved for ``tag`` is durable."""
        return True

    # -- stateless helpers ---------------------------------------------------------------------------------------------
    @abc.abstractmethod
    def load(self, path: str, map_location=None):
        ...

    def makedirs(self, path, exist_ok=False):
        os.makedirs(path, exist_ok=exist_ok)
