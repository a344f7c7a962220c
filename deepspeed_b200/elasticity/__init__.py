from .elasticity import (compute_elastic_config, elasticity_enabled, ensure_immutable_elastic_config,  # noqa: F401
                         highly_composite_numbers)
from .config import ElasticityConfig, ElasticityConfigError, ElasticityError, ElasticityIncompatibleWorldSize  # noqa: F401
from .elastic_agent import DSElasticAgent  # noqa: F401
