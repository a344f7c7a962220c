"""``"timers"`` config section (reference ``utils/config.py``)."""
from deepspeed_b200.runtime.config_utils import DeepSpeedConfigModel

TIMERS = "timers"
TIMERS_THROUGHPUT = "throughput"
TIMERS_FORMAT = '"timers": {"throughput": {"enabled": true, "synchronized": true}}'


class DeepSpeedThroughputTimerConfig(DeepSpeedConfigModel):
    enabled: bool = True
    synchronized: bool = True  # device sync around the measurement (accurate) vs host clock only (cheap)


def get_timers_config(param_dict):
    section = (param_dict or {}).get(TIMERS, {}).get(TIMERS_THROUGHPUT, {})
    return DeepSpeedThroughputTimerConfig(**section)
