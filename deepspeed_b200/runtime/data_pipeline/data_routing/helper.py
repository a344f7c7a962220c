"""Wrap / unwrap model layers with random-LTD (reference ``data_routing/helper.py``)."""
from .basic_layer import RandomLayerTokenDrop


def convert_to_random_ltd(model, convert_type):
    """Replace every sub-module of class ``convert_type`` by a ``RandomLayerTokenDrop`` around it."""
    if hasattr(model, "module"):
        model = model.module
    for name, mod in list(model.named_modules()):
        for cname, child in list(mod.named_children()):
            if isinstance(child, convert_type) and not isinstance(child, RandomLayerTokenDrop):
                setattr(mod, cname, RandomLayerTokenDrop(child))
    return model


def save_without_random_ltd(model):
    """State dict with the wrapper's ``random_ltd_layer.`` prefix stripped (checkpoint compatibility)."""
    if hasattr(model, "module"):
        model = model.module
    return remove_random_ltd_state_dict(model.state_dict())


def remove_random_ltd_state_dict(state_dict):
    return {k.replace(".random_ltd_layer", ""): v for k, v in state_dict.items()}
