from deepspeed_b200.utils.torch import required_torch_version


def is_torch_elastic_compatible():
    """torch.distributed.elastic (the agent the elastic launcher builds on) exists since torch 1.11."""
    return required_torch_version(min_version=1.11)
