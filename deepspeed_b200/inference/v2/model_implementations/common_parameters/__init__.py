from .attn_output_parameters import *  # noqa: F401,F403
from .embedding_parameters import *  # noqa: F401,F403
from .invfreq_parameters import *  # noqa: F401,F403
from .mlp_parameters import *  # noqa: F401,F403
from .moe_parameters import *  # noqa: F401,F403
from .norm_parameters import *  # noqa: F401,F403
from .qkv_parameters import *  # noqa: F401,F403
from .unembed_parameters import *  # noqa: F401,F403
