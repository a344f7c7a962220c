"""``"activation_checkpointing"`` config section (reference ``runtime/activation_checkpointing/config.py``)."""
from deepspeed_b200.runtime.config import ActivationCheckpointingConfig

ACT_CHKPT = "activation_checkpointing"
ACT_CHKPT_PARTITION_ACTIVATIONS = "partition_activations"
ACT_CHKPT_NUMBER_CHECKPOINTS = "number_checkpoints"
ACT_CHKPT_CONTIGUOUS_MEMORY_OPTIMIZATION = "contiguous_memory_optimization"
ACT_CHKPT_SYNCHRONIZE_CHECKPOINT_BOUNDARY = "synchronize_checkpoint_boundary"
ACT_CHKPT_PROFILE = "profile"
ACT_CHKPT_CPU_CHECKPOINTING = "cpu_checkpointing"
ACT_CHKPT_DEFAULT = ActivationCheckpointingConfig().model_dump()


class DeepSpeedActivationCheckpointingConfig(ActivationCheckpointingConfig):

    def __init__(self, param_dict=None, **kw):
        super().__init__(**{**(param_dict or {}).get(ACT_CHKPT, {}), **kw})
