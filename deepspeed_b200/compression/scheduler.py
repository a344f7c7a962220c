"""Turn techniques on at their ``schedule_offset`` (reference ``compression/scheduler.py:12``)."""
from . import constants as C
from .compress import get_module_name
from .config import get_compression_config
from .helper import recursive_getattr

_FLAG = {C.WEIGHT_QUANTIZATION: "weight_quantization_enabled", C.ACTIVATION_QUANTIZATION: "activation_quantization_enabled",
         C.SPARSE_PRUNING: "sparse_pruning_enabled", C.ROW_PRUNING: "row_pruning_enabled",
         C.HEAD_PRUNING: "head_pruning_enabled", C.CHANNEL_PRUNING: "channel_pruning_enabled"}


class compression_scheduler:

    def __init__(self, model, compression_config):
        self.model = model
        self.compression_config = compression_config
        self.training_steps = 0
        self.weight_quantization_enabled = False
        self.verbose = {t: False for t in C.TECHNIQUES}
        self.different_compression_methods = {}
        for method, mc in compression_config.items():
            if method == C.LAYER_REDUCTION:
                continue
            shared = mc[C.SHARED_PARAMETERS]
            entry = {C.TECHNIQUE_ENABLED: shared[C.TECHNIQUE_ENABLED], C.SHARED_PARAMETERS: shared, C.DIFFERENT_GROUPS: []}
            seen = []
            for gname, g in mc[C.DIFFERENT_GROUPS].items():
                names = []
                for kw in g[C.DIFFERENT_GROUPS_MODULE_SCOPE]:
                    found, seen = get_module_name(gname, model, kw, seen, verbose=False)
                    names.extend(found)
                if names:
                    entry[C.DIFFERENT_GROUPS].append([gname, names, g[C.DIFFERENT_GROUPS_PARAMETERS]])
            self.different_compression_methods[method] = entry

    def _check(self, method):
        e = self.different_compression_methods.get(method)
        if not e or not e[C.TECHNIQUE_ENABLED]:
            return
        if self.training_steps >= e[C.SHARED_PARAMETERS][C.TECHNIQUE_SCHEDULE_OFFSET]:
            for _, names, _ in e[C.DIFFERENT_GROUPS]:
                for n in names:
                    setattr(recursive_getattr(self.model, n), _FLAG[method], True)
            if not self.verbose[method]:
                self.verbose[method] = True
            if method == C.WEIGHT_QUANTIZATION:
                self.weight_quantization_enabled = True

    def step(self, step_zero_check=False):
        if not step_zero_check:
            self.training_steps += 1
        for m in C.TECHNIQUES:
            self._check(m)
