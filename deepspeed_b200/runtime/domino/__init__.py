from .transformer import DominoTransformerLayer, DominoTransformer  # noqa: F401
