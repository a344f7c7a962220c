"""Availability checks for the optional monitor back-ends (reference ``monitor/utils.py``)."""
import importlib

from packaging import version as _v


def _require(module, hint, min_version=None):
    """Import ``module`` or print the install hint and re-raise."""
    try:
        mod = importlib.import_module(module)
        if min_version is not None and _v.parse(mod.__version__) < _v.Version(min_version):
            raise ImportError(f"`{module}` must have at least version {min_version}")
        return mod
    except ImportError:
        print(hint)
        raise


def check_tb_availability():
    return _require("tensorboard", "If you want to use tensorboard logging, please `pip install tensorboard`")


def check_wandb_availability():
    return _require("wandb", "If you want to use wandb logging, please `pip install wandb` and follow https://docs.wandb.ai/quickstart")


def check_comet_availability():
    return _require("comet_ml", 'If you want to use comet logging, please `pip install "comet_ml>=3.41.0"`', "3.41.0")
