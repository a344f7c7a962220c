"""DeepNVMe tools: ``ds_io`` (one read/write benchmark) and ``ds_nvme_tune`` (sweep -> recommended aio config).
Reference: ``deepspeed/nvme/{ds_aio_*.py, perf_run_sweep.py, perf_generate_param.py}`` and ``bin/ds_io|ds_nvme_tune``."""
from .perf import run_io_benchmark, ds_io_main  # noqa: F401
from .sweep import run_sweep, generate_aio_param, sweep_main  # noqa: F401
