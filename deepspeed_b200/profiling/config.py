from deepspeed_b200.runtime.config import FlopsProfilerConfig as DeepSpeedFlopsProfilerConfig  # noqa: F401
