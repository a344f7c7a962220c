"""Attach the high-precision accessors to low-precision parameters (reference ``utils/mixed_precision_linkage.py``).

In this framework the lp -> hp mapping is a property of the ZeRO unit plan (``runtime/zero/units.py``), so linking is
just binding the accessor methods and caching the fragment list; there are no per-group flat partitions to walk.
"""
import types

from .tensor_fragment import (get_hp_fragment_mapping, safe_get_full_fp32_param, safe_get_full_grad,
                              safe_set_full_fp32_param, safe_set_full_grad)


def link_hp_params(lp_param_list, flat_hp_partition=None, gradient_dict=None, offload_gradient_dict=None, use_offload=False,
                   param_group_index=0, partition_start=None, partition_size=None, dp_group=None):
    for idx, lp in enumerate(lp_param_list):
        lp._dp_group = dp_group
        lp.get_full_hp_param = types.MethodType(lambda self, optim_state_key=None: safe_get_full_fp32_param(self), lp)
        lp.get_full_hp_grad = types.MethodType(lambda self: safe_get_full_grad(self), lp)
        lp.set_full_hp_param = types.MethodType(lambda self, value, optim_state_key=None: safe_set_full_fp32_param(self, value), lp)
        lp.set_full_hp_grad = types.MethodType(lambda self, value: safe_set_full_grad(self, value), lp)
        frags = get_hp_fragment_mapping(lp)
        lp._hp_mapping = frags or None
        lp._index_in_param_group = idx


def lazy_init_hp_params_optimizer_state(lp_param_list, flat_hp_partition=None, optimizer_state=None):
    """Optimizer moments live in flat arenas allocated with the optimizer, so there is nothing to create lazily; kept so
    code written against the reference keeps working."""
    return None
