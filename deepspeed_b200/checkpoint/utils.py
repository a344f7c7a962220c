"""Checkpoint file-name helpers (reference ``checkpoint/utils.py``)."""
import os

import torch

from .constants import (LAYER_FILE_PREFIX, MODEL_FILE_PREFIX, MODEL_FILE_SUFFIX, OPTIM_FILE_SUFFIX, ZERO_FILE_PREFIX)


def get_model_ckpt_name_for_rank(base_folder, mp_rank_str):
    return os.path.join(base_folder, MODEL_FILE_PREFIX + mp_rank_str + MODEL_FILE_SUFFIX)


def get_zero_ckpt_name_for_rank(base_folder, dp_rank, mp_rank):
    return os.path.join(base_folder, f"{ZERO_FILE_PREFIX}{dp_rank}_{MODEL_FILE_PREFIX}{mp_rank:02d}{OPTIM_FILE_SUFFIX}")


def get_layer_ckpt_name_for_rank(base_folder, layer_id, tp_rank):
    return os.path.join(base_folder, f"{LAYER_FILE_PREFIX}{layer_id}-model_{tp_rank:02d}{MODEL_FILE_SUFFIX}")


def clone_tensors_for_torch_save(item, device=torch.device("cpu")):
    """Deep copy with tensors detached from their (possibly huge, flat) storage so ``torch.save`` writes only
    the viewed bytes (reference ``checkpoint/utils.py:43``)."""
    if torch.is_tensor(item):
        return item.detach().clone().to(device)
    if isinstance(item, (list, tuple)):
        return type(item)(clone_tensors_for_torch_save(v, device) for v in item)
    if isinstance(item, dict):
        return type(item)({k: clone_tensors_for_torch_save(v, device) for k, v in item.items()})
    return item
