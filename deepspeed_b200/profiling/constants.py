"""``"flops_profiler"`` config keys (reference ``profiling/constants.py``).

Names are declared section by section through ``_declare`` (later sections may refer to earlier names)."""


def _declare(**names):
    globals().update(names)
    return names


_declare(
    FLOPS_PROFILER="flops_profiler",
    FLOPS_PROFILER_ENABLED="enabled",
    FLOPS_PROFILER_ENABLED_DEFAULT=False,
    FLOPS_PROFILER_RECOMPUTE_FWD_FACTOR="recompute_fwd_factor",
    FLOPS_PROFILER_RECOMPUTE_FWD_FACTOR_DEFAULT=0.0,
    FLOPS_PROFILER_PROFILE_STEP="profile_step",
    FLOPS_PROFILER_PROFILE_STEP_DEFAULT=1,
    FLOPS_PROFILER_MODULE_DEPTH="module_depth",
    FLOPS_PROFILER_MODULE_DEPTH_DEFAULT=-1,
    FLOPS_PROFILER_TOP_MODULES="top_modules",
    FLOPS_PROFILER_TOP_MODULES_DEFAULT=1,
    FLOPS_PROFILER_DETAILED="detailed",
    FLOPS_PROFILER_DETAILED_DEFAULT=True,
    FLOPS_PROFILER_OUTPUT_FILE="output_file",
    FLOPS_PROFILER_OUTPUT_FILE_DEFAULT=None,
)
