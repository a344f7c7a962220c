"""Model zoo used by the benchmarks, tests and inference engines."""
from .llama import LlamaConfig, LlamaForCausalLM, LlamaModel, llama_config  # noqa: F401
