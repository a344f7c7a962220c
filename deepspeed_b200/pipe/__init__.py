from ..runtime.pipe import LayerSpec, PipelineModule, ProcessTopology, TiedLayerSpec  # noqa: F401
