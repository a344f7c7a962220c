from .module import LayerSpec, PipelineModule, TiedLayerSpec  # noqa: F401
from .topology import ProcessTopology  # noqa: F401
