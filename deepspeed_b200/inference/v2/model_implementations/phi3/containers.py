from .container import Phi3NonTransformerContainer, Phi3TransformerContainer  # noqa: F401  (reference file name: containers.py)
