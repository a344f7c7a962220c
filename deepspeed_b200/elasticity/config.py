"""Elasticity config object (reference ``elasticity/config.py``)."""
import json

from . import constants as C


class ElasticityError(Exception):
    """Base exception for all elasticity related errors."""


class ElasticityConfigError(ElasticityError):
    """Elasticity configuration error."""


class ElasticityIncompatibleWorldSize(ElasticityError):
    """Attempting to run a world size that is incompatible with a given elastic config."""


class ElasticityConfig:
    """
    ::

        "elasticity": {"enabled": true, "max_train_batch_size": 2000, "micro_batch_sizes": [2,4,6],
                       "min_gpus": 1, "max_gpus": 10000, "min_time": 20, "ignore_non_elastic_batch_info": false,
                       "version": 0.2, "num_gpus_per_node": 8, "model_parallel_size": 1}
    """

    def __init__(self, param_dict):
        g = param_dict.get
        self.enabled = g(C.ENABLED, C.ENABLED_DEFAULT)
        if self.enabled:
            for key in (C.MAX_ACCEPTABLE_BATCH_SIZE, C.MICRO_BATCHES):
                if key not in param_dict:
                    raise ElasticityConfigError(f"Elasticity config missing {key}")
        self.max_acceptable_batch_size = g(C.MAX_ACCEPTABLE_BATCH_SIZE, C.MAX_ACCEPTABLE_BATCH_SIZE_DEFAULT)
        self.micro_batches = g(C.MICRO_BATCHES, C.MICRO_BATCHES_DEFAULT)
        if not isinstance(self.micro_batches, list):
            raise ElasticityConfigError(f"Elasticity expected value of {C.MICRO_BATCHES} to be a list of micro "
                                        f"batches, instead is: {type(self.micro_batches)}, containing: "
                                        f"{self.micro_batches}")
        if not all(isinstance(m, int) for m in self.micro_batches):
            raise ElasticityConfigError(f"Elasticity expected {C.MICRO_BATCHES} to only contain a list of integers, "
                                        f"instead contains: {self.micro_batches}")
        if not all(m > 0 for m in self.micro_batches):
            raise ElasticityConfigError(f"Elasticity expected {C.MICRO_BATCHES} to only contain positive integers, "
                                        f"instead contains: {self.micro_batches}")
        self.min_gpus = g(C.MIN_GPUS, C.MIN_GPUS_DEFAULT)
        self.max_gpus = g(C.MAX_GPUS, C.MAX_GPUS_DEFAULT)
        if self.min_gpus < 1 or self.max_gpus < 1:
            raise ElasticityConfigError(f"Elasticity min/max gpus must be > 0, given min_gpus: {self.min_gpus}, "
                                        f"max_gpus: {self.max_gpus}")
        if self.max_gpus < self.min_gpus:
            raise ElasticityConfigError(f"Elasticity min_gpus cannot be greater than max_gpus, given min_gpus: "
                                        f"{self.min_gpus}, max_gpus: {self.max_gpus}")
        self.model_parallel_size = g(C.MODEL_PARALLEL_SIZE, C.MODEL_PARALLEL_SIZE_DEFAULT)
        if self.model_parallel_size < 1:
            raise ElasticityConfigError(f"Model-Parallel size cannot be less than 1, given {self.model_parallel_size}")
        self.num_gpus_per_node = g(C.NUM_GPUS_PER_NODE, C.NUM_GPUS_PER_NODE_DEFAULT)
        if self.num_gpus_per_node < 1:
            raise ElasticityConfigError(f"Number of GPUs per node cannot be less than 1, given {self.num_gpus_per_node}")
        self.min_time = g(C.MIN_TIME, C.MIN_TIME_DEFAULT)
        if self.min_time < 0:
            raise ElasticityConfigError(f"Elasticity min time needs to be >= 0: given {self.min_time}")
        self.version = g(C.VERSION, C.VERSION_DEFAULT)
        self.prefer_larger_batch_size = g(C.PREFER_LARGER_BATCH, C.PREFER_LARGER_BATCH_DEFAULT)
        self.ignore_non_elastic_batch_info = g(C.IGNORE_NON_ELASTIC_BATCH_INFO, C.IGNORE_NON_ELASTIC_BATCH_INFO_DEFAULT)

    def repr(self):
        return self.__dict__

    def __repr__(self):
        return json.dumps(self.__dict__, sort_keys=True, indent=4)
