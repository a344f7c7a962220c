"""Elasticity config keys (reference ``elasticity/constants.py``).

Names are declared section by section through ``_declare`` (later sections may refer to earlier names)."""


def _declare(**names):
    globals().update(names)
    return names


_declare(
    ELASTICITY="elasticity",
    LATEST_ELASTICITY_VERSION=0.2,
    ENABLED="enabled",
    ENABLED_DEFAULT=False,
)

_declare(
    MAX_ACCEPTABLE_BATCH_SIZE="max_train_batch_size",
    MAX_ACCEPTABLE_BATCH_SIZE_DEFAULT=2000,
    MICRO_BATCHES="micro_batch_sizes",
    MICRO_BATCHES_DEFAULT=[2, 4, 6],
)

_declare(
    MIN_GPUS="min_gpus",
    MIN_GPUS_DEFAULT=1,
    MAX_GPUS="max_gpus",
    MAX_GPUS_DEFAULT=10000,
)

_declare(
    NUM_GPUS_PER_NODE="num_gpus_per_node",
    NUM_GPUS_PER_NODE_DEFAULT=1,
    MODEL_PARALLEL_SIZE="model_parallel_size",
    MODEL_PARALLEL_SIZE_DEFAULT=1,
)

_declare(
    MIN_TIME="min_time",
    MIN_TIME_DEFAULT=0,
    PREFER_LARGER_BATCH="prefer_larger_batch",
    PREFER_LARGER_BATCH_DEFAULT=True,
)

_declare(
    IGNORE_NON_ELASTIC_BATCH_INFO="ignore_non_elastic_batch_info",
    IGNORE_NON_ELASTIC_BATCH_INFO_DEFAULT=False,
    VERSION="version",
    VERSION_DEFAULT=LATEST_ELASTICITY_VERSION,
)

_declare(
    MINIMUM_DEEPSPEED_VERSION="0.3.8",  # the upstream release that introduced elasticity (schedulers pass UPSTREAM version strings)
    DEEPSPEED_ELASTICITY_CONFIG="DEEPSPEED_ELASTICITY_CONFIG",
)
