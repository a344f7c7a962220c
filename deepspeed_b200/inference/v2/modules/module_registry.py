"""Registries mapping implementation names to module classes (reference ``inference/v2/modules/module_registry.py``)."""
from typing import Any, Dict, Type

from deepspeed_b200.runtime.config_utils import DeepSpeedConfigModel

from .ds_module import DSModuleBase


class ConfigBundle(DeepSpeedConfigModel):
    """What a heuristic hands to a registry: implementation name + module config (+ implementation-specific knobs)."""
    name: str
    config: Any = None
    implementation_config: Dict[str, Any] = {}


class DSModuleRegistryBase:
    """One subclass per module kind; ``registry`` is per subclass."""
    registry: Dict[str, Type[DSModuleBase]] = None

    def __init_subclass__(cls, **kw):
        super().__init_subclass__(**kw)
        cls.registry = {}

    @classmethod
    def associated_class(cls) -> Type[DSModuleBase]:
        raise NotImplementedError

    @classmethod
    def instantiate_config(cls, config_bundle: ConfigBundle) -> DSModuleBase:
        if config_bundle.name not in cls.registry:
            raise KeyError(f"Unknown DSModule: {config_bundle.name}, cls.registry={sorted(cls.registry)}")
        target = cls.registry[config_bundle.name]
        if not target.supports_config(config_bundle.config):
            raise ValueError(f"Config {config_bundle.config} is not supported by {target}")
        return target(config_bundle.config, config_bundle.implementation_config)

    @classmethod
    def register_module(cls, child_class: Type[DSModuleBase]) -> Type[DSModuleBase]:
        if not issubclass(child_class, cls.associated_class()):
            raise TypeError(f"Can only register subclasses of {cls.associated_class()}, got {child_class}")
        cls.registry[child_class.name()] = child_class
        return child_class
