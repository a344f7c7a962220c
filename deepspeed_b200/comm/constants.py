"""Backend names + comms-logger config keys (reference ``comm/constants.py``).  Only NCCL (device) and gloo (host tier
and tests) are live backends here; the other names are kept so configs that mention them fail with a clear message.

Names are declared section by section through ``_declare`` (later sections may refer to earlier names)."""


def _declare(**names):
    globals().update(names)
    return names


_declare(
    NCCL_BACKEND="nccl",
    GLOO_BACKEND="gloo",
    MPI_BACKEND="mpi",
    CCL_BACKEND="ccl",
)

_declare(
    SCCL_BACKEND="sccl",
    HCCL_BACKEND="hccl",
    SUPPORTED_BACKENDS=(NCCL_BACKEND, GLOO_BACKEND),
    DEFAULT_AML_MASTER_PORT="54965",
    DEFAULT_AML_NCCL_SOCKET_IFNAME="^docker0,lo",
)

_declare(
    COMMS_LOGGER="comms_logger",
    COMMS_LOGGER_ENABLED="enabled",
    COMMS_LOGGER_ENABLED_DEFAULT=False,
    COMMS_LOGGER_VERBOSE="verbose",
    COMMS_LOGGER_VERBOSE_DEFAULT=False,
    COMMS_LOGGER_PROF_ALL="prof_all",
    COMMS_LOGGER_PROF_ALL_DEFAULT=True,
    COMMS_LOGGER_DEBUG="debug",
    COMMS_LOGGER_DEBUG_DEFAULT=False,
    COMMS_LOGGER_PROF_OPS="prof_ops",
    COMMS_LOGGER_PROF_OPS_DEFAULT=[],
    COMMS_LOGGER_FORMAT='"comms_logger": {"enabled": true, "verbose": false, "prof_all": true, "debug": false, '
                       '"prof_ops": ["all_reduce", "custom_all_reduce_name"]}',
)
