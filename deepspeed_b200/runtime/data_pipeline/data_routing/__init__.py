from .basic_layer import RandomLayerTokenDrop  # noqa: F401
from .scheduler import RandomLTDScheduler  # noqa: F401
from .helper import convert_to_random_ltd, save_without_random_ltd  # noqa: F401
