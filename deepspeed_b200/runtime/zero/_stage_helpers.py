"""Small module-level helpers the reference defines next to its ZeRO optimizers (``stage_1_and_2.py:45-95``,
``stage3.py:40-75``); shared by the two compatibility modules."""
import math

import torch

from deepspeed_b200 import comm as dist

# The reference's module-level debug switch (``stage_1_and_2.py:36``).  Setting it to True before ``initialize()`` turns on
# the run-time cross-check of every in-kernel NVLink collective against NCCL for the whole run (the config knob
# ``zero_optimization.b200_verify_collectives: N`` does the same for the first N steps); see ``sharded.py``.
pg_correctness_test = False
OPTIMIZER_ALLGATHER_TIMER, OPTIMIZER_GRADIENTS_TIMER, OPTIMIZER_STEP_TIMER = "optimizer_allgather", "optimizer_gradients", \
    "optimizer_step"
OPTIMIZER_TIMERS = [OPTIMIZER_ALLGATHER_TIMER, OPTIMIZER_GRADIENTS_TIMER, OPTIMIZER_STEP_TIMER]
OPTIMIZER_SWAP_IN_STATE_TIMER, INIT_OPTIMIZER_TIMER = "optimizer_swap_in_state", "init_optimizer_state"
OPTIMIZER_SWAP_OUT_STATE_TIMER = "optimizer_swap_out_state"
INITIAL_MICRO_STEP_ID = -1


def input(msg):  # noqa: A001  (debug hook of the reference: a no-op that shadows the builtin on purpose)
    return


def split_half_float_double(tensors):
    """Bucket tensors by dtype in the order fp16, fp32, fp64, bf16 (empty buckets dropped)."""
    order = (torch.float16, torch.float32, torch.float64, torch.bfloat16)
    buckets = [[t for t in tensors if t.dtype == dt] for dt in order]
    return [b for b in buckets if b]


def isclose(a, b, rtol=1e-09, atol=0.0):
    return abs(a - b) <= max(rtol * max(abs(a), abs(b)), atol)


def lcm(x, y):
    return x * y // math.gcd(x, y)


def get_alignment_padding(tensor_list, alignment):
    """Elements to append so the flattened list is a multiple of ``alignment``."""
    return -sum(t.numel() for t in tensor_list) % alignment


def print_rank_msg(msg):
    print(f"rank {dist.get_rank() if dist.is_initialized() else 0} - {msg}")


def print_rank_0(message, debug=False, force=False):
    if (debug or force) and (not dist.is_initialized() or dist.get_rank() == 0):
        print(message)


def move_to_cpu(tensor_list):
    for t in tensor_list:
        t.data = t.data.cpu()


def model_to_params(model):
    """``(total elements, trainable parameter list)``."""
    ps = [p for p in model.parameters() if p.requires_grad]
    return sum(getattr(p, "ds_numel", p.numel()) for p in ps), ps
