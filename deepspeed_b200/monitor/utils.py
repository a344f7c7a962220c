"""Availability checks for the optional monitor back-ends (reference ``monitor/utils.py``)."""
from packaging import version as _v


def check_tb_availability():
    try:
        import tensorboard  # noqa: F401
    except ImportError:
        print("If you want to use tensorboard logging, please `pip install tensorboard`")
        raise


def check_wandb_availability():
    try:
        import wandb  # noqa: F401
    except ImportError:
        print("If you want to use wandb logging, please `pip install wandb` and follow https://docs.wandb.ai/quickstart")
        raise


def check_comet_availability():
    try:
        import comet_ml
        if _v.parse(comet_ml.__version__) < _v.Version("3.41.0"):
            raise ImportError("`comet_ml` must have at least version 3.41.0")
    except ImportError:
        print('If you want to use comet logging, please `pip install "comet_ml>=3.41.0"`')
        raise
