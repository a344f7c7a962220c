"""``"nebula"`` config keys (reference ``nebula/constants.py``).

Names are declared section by section through ``_declare`` (later sections may refer to earlier names)."""


def _declare(**names):
    globals().update(names)
    return names


_declare(
    NEBULA="nebula",
    NEBULA_ENABLED="enabled",
    NEBULA_ENABLED_DEFAULT=False,
    NEBULA_ENABLE_NEBULA_LOAD="enable_nebula_load",
    NEBULA_ENABLE_NEBULA_LOAD_DEFAULT=True,
    NEBULA_LOAD_PATH="nebula_load_path",
    NEBULA_LOAD_PATH_DEFAULT=None,
    NEBULA_PERSISTENT_STORAGE_PATH="persistent_storage_path",
    NEBULA_PERSISTENT_STORAGE_PATH_DEFAULT=None,
    NEBULA_PERSISTENT_TIME_INTERVAL="persistent_time_interval",
    NEBULA_PERSISTENT_TIME_INTERVAL_DEFAULT=100,
    NEBULA_NUM_OF_VERSION_IN_RETENTION="num_of_version_in_retention",
    NEBULA_NUM_OF_VERSION_IN_RETENTION_DEFAULT=2,
    NEBULA_FORMAT='"nebula": {"enabled": true, "persistent_storage_path": "/foo/bar", "persistent_time_interval": 100, '
                 '"num_of_version_in_retention": 2, "enable_nebula_load": true}',
)

_declare(
    NEBULA_EXPORT_ENVS=["NEBULA_PERSISTENT_STORAGE_PATH", "NEBULA_PERSISTENT_TIME_INTERVAL", "NEBULA_MEMORY_BUFFER_SIZE",
                      "MASTER_HOST", "LOCAL_HOST"],
)
