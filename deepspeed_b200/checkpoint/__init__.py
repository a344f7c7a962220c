from .constants import *  # noqa: F401,F403
from .deepspeed_checkpoint import DeepSpeedCheckpoint  # noqa: F401
from .universal_checkpoint import load_universal_into_engine, load_hp_checkpoint_state  # noqa: F401
from .ds_to_universal import convert_to_universal  # noqa: F401
from .utils import (get_model_ckpt_name_for_rank, get_zero_ckpt_name_for_rank, get_layer_ckpt_name_for_rank,  # noqa: F401
                    clone_tensors_for_torch_save)
from .reshape_utils import merge_state, partition_data, get_zero_files  # noqa: F401,E402
from .reshape_meg_2d import meg_2d_parallel_map, reshape_meg_2d_parallel, get_mpu_ranks  # noqa: F401,E402
from .reshape_3d_utils import model_3d_desc, get_model_3d_descriptor  # noqa: F401,E402
from .zero_checkpoint import ZeROCheckpoint  # noqa: F401,E402

from .universal_checkpoint import SubparamShape, enable_universal_checkpoint  # noqa: F401,E402
