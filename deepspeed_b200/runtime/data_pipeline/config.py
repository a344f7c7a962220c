"""Parse the ``data_efficiency`` (and legacy ``curriculum_learning``) config blocks with defaults
(reference ``runtime/data_pipeline/config.py``)."""
import copy

from . import constants as C


def get_data_efficiency_config(param_dict):
    sub = copy.deepcopy(param_dict.get(C.DATA_EFFICIENCY, {}))
    out = {
        C.DATA_EFFICIENCY_ENABLED: sub.get(C.DATA_EFFICIENCY_ENABLED, False),
        C.DATA_EFFICIENCY_SEED: sub.get(C.DATA_EFFICIENCY_SEED, C.DATA_EFFICIENCY_SEED_DEFAULT),
    }
    ds = copy.deepcopy(sub.get(C.DATA_SAMPLING, {}))
    ds.setdefault(C.DATA_SAMPLING_ENABLED, False)
    ds.setdefault(C.DATA_SAMPLING_NUM_EPOCHS, C.DATA_SAMPLING_NUM_EPOCHS_DEFAULT)
    ds.setdefault(C.DATA_SAMPLING_NUM_WORKERS, C.DATA_SAMPLING_NUM_WORKERS_DEFAULT)
    cl = ds.setdefault(C.CURRICULUM_LEARNING, {})
    cl.setdefault(C.CURRICULUM_LEARNING_ENABLED, False)
    out[C.DATA_SAMPLING] = ds
    dr = copy.deepcopy(sub.get(C.DATA_ROUTING, {}))
    dr.setdefault(C.DATA_ROUTING_ENABLED, False)
    ltd = dr.setdefault(C.RANDOM_LTD, {})
    ltd.setdefault(C.RANDOM_LTD_ENABLED, False)
    out[C.DATA_ROUTING] = dr
    return out


def get_curriculum_enabled_legacy(param_dict):
    return param_dict.get(C.CURRICULUM_LEARNING, {}).get(C.CURRICULUM_LEARNING_ENABLED, False)


def get_curriculum_params_legacy(param_dict):
    d = copy.copy(param_dict.get(C.CURRICULUM_LEARNING, {}))
    d.pop(C.CURRICULUM_LEARNING_ENABLED, None)
    return d or False


# ---- section accessors (reference ``runtime/data_pipeline/config.py:get_*``); ``param_dict`` is the section's parent ----
def _full(param_dict):
    return get_data_efficiency_config({C.DATA_EFFICIENCY: param_dict}) if C.DATA_EFFICIENCY not in param_dict else \
        get_data_efficiency_config(param_dict)


def get_data_efficiency_enabled(param_dict):
    return _full(param_dict)[C.DATA_EFFICIENCY_ENABLED]


def get_data_efficiency_seed(param_dict):
    return _full(param_dict)[C.DATA_EFFICIENCY_SEED]


def get_data_sampling(param_dict):
    return _full(param_dict)[C.DATA_SAMPLING]


def get_data_sampling_enabled(param_dict):
    return get_data_sampling(param_dict)[C.DATA_SAMPLING_ENABLED]


def get_data_sampling_num_epochs(param_dict):
    return get_data_sampling(param_dict)[C.DATA_SAMPLING_NUM_EPOCHS]


def get_data_sampling_num_workers(param_dict):
    return get_data_sampling(param_dict)[C.DATA_SAMPLING_NUM_WORKERS]


def get_curriculum_learning(param_dict):
    return get_data_sampling(param_dict)[C.CURRICULUM_LEARNING]


def get_curriculum_learning_enabled(param_dict):
    return get_curriculum_learning(param_dict)[C.CURRICULUM_LEARNING_ENABLED]


def get_curriculum_learning_params(param_dict):
    d = dict(get_curriculum_learning(param_dict))
    d.pop(C.CURRICULUM_LEARNING_ENABLED, None)
    return d or False


def get_data_routing(param_dict):
    return _full(param_dict)[C.DATA_ROUTING]


def get_data_routing_enabled(param_dict):
    return get_data_routing(param_dict)[C.DATA_ROUTING_ENABLED]


def get_random_ltd(param_dict):
    return get_data_routing(param_dict)[C.RANDOM_LTD]


def get_random_ltd_enabled(param_dict):
    return get_random_ltd(param_dict)[C.RANDOM_LTD_ENABLED]


def get_random_ltd_params(param_dict):
    d = dict(get_random_ltd(param_dict))
    d.pop(C.RANDOM_LTD_ENABLED, None)
    return d or False
