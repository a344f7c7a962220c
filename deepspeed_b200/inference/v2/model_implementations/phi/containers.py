from .container import PhiNonTransformerContainer, PhiTransformerContainer  # noqa: F401  (reference file name: containers.py)
