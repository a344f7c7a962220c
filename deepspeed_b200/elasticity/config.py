"""Elasticity config object (reference ``elasticity/config.py``).

The ``"elasticity"`` section is validated by a table of (key, default, check) rules; every violation is reported as an
``ElasticityConfigError`` naming the offending key::

    "elasticity": {"enabled": true, "max_train_batch_size": 2000, "micro_batch_sizes": [2, 4, 6], "min_gpus": 1,
                   "max_gpus": 10000, "min_time": 20, "ignore_non_elastic_batch_info": false, "version": 0.2,
                   "num_gpus_per_node": 8, "model_parallel_size": 1}
"""
import json

from . import constants as C


class ElasticityError(Exception):
    """Root of the elasticity exceptions."""


class ElasticityConfigError(ElasticityError):
    """The ``elasticity`` config section is malformed or inconsistent."""


class ElasticityIncompatibleWorldSize(ElasticityError):
    """The job was started on a world size the elastic config cannot serve."""


def _is_pos_int_list(v):
    return isinstance(v, list) and all(isinstance(m, int) and not isinstance(m, bool) and m > 0 for m in v)


# attribute, config key, default, predicate, requirement in words
_RULES = (
    ("max_acceptable_batch_size", C.MAX_ACCEPTABLE_BATCH_SIZE, C.MAX_ACCEPTABLE_BATCH_SIZE_DEFAULT, None, ""),
    ("micro_batches", C.MICRO_BATCHES, C.MICRO_BATCHES_DEFAULT, _is_pos_int_list, "a list of positive integers"),
    ("min_gpus", C.MIN_GPUS, C.MIN_GPUS_DEFAULT, lambda v: v >= 1, ">= 1"),
    ("max_gpus", C.MAX_GPUS, C.MAX_GPUS_DEFAULT, lambda v: v >= 1, ">= 1"),
    ("model_parallel_size", C.MODEL_PARALLEL_SIZE, C.MODEL_PARALLEL_SIZE_DEFAULT, lambda v: v >= 1, ">= 1"),
    ("num_gpus_per_node", C.NUM_GPUS_PER_NODE, C.NUM_GPUS_PER_NODE_DEFAULT, lambda v: v >= 1, ">= 1"),
    ("min_time", C.MIN_TIME, C.MIN_TIME_DEFAULT, lambda v: v >= 0, ">= 0"),
    ("version", C.VERSION, C.VERSION_DEFAULT, None, ""),
    ("prefer_larger_batch_size", C.PREFER_LARGER_BATCH, C.PREFER_LARGER_BATCH_DEFAULT, None, ""),
    ("ignore_non_elastic_batch_info", C.IGNORE_NON_ELASTIC_BATCH_INFO, C.IGNORE_NON_ELASTIC_BATCH_INFO_DEFAULT, None, ""),
)
_REQUIRED_WHEN_ENABLED = (C.MAX_ACCEPTABLE_BATCH_SIZE, C.MICRO_BATCHES)


class ElasticityConfig:

    def __init__(self, param_dict):
        self.enabled = param_dict.get(C.ENABLED, C.ENABLED_DEFAULT)
        missing = [k for k in _REQUIRED_WHEN_ENABLED if self.enabled and k not in param_dict]
        if missing:
            raise ElasticityConfigError(f"elasticity is enabled but {missing} {'is' if len(missing) == 1 else 'are'} not set")
        for attr, key, default, ok, words in _RULES:
            value = param_dict.get(key, default)
            if ok is not None and not ok(value):
                raise ElasticityConfigError(f"elasticity.{key} must be {words}, got {value!r}")
            setattr(self, attr, value)
        if self.max_gpus < self.min_gpus:
            raise ElasticityConfigError(f"elasticity.{C.MIN_GPUS} ({self.min_gpus}) exceeds elasticity.{C.MAX_GPUS} ({self.max_gpus})")

    def repr(self):
        return self.__dict__

    def __repr__(self):
        return json.dumps(self.__dict__, sort_keys=True, indent=4)
