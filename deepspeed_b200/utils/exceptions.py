class DeprecatedException(Exception):
    """Raised when a removed config key / API is used (reference ``utils/exceptions.py``)."""
