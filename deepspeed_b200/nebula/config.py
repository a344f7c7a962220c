"""``DeepSpeedNebulaConfig`` (reference ``nebula/config.py``) over the pydantic section in ``runtime/config.py``."""
from deepspeed_b200.nebula.constants import NEBULA, NEBULA_LOAD_PATH
from deepspeed_b200.runtime.config import NebulaConfig


class DeepSpeedNebulaConfig(NebulaConfig):

    def __init__(self, param_dict=None, **kw):
        section = dict((param_dict or {}).get(NEBULA, {}))
        if NEBULA_LOAD_PATH in section:
            section["load_path"] = section.pop(NEBULA_LOAD_PATH)
        super().__init__(**{**section, **kw})
