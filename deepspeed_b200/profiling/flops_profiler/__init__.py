from .profiler import (FlopsProfiler, get_model_profile, flops_to_string, macs_to_string, params_to_string,  # noqa: F401
                       duration_to_string, number_to_string, bytes_to_string)
