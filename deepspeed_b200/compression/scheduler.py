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
        self.make_init()

    def make_init(self):
        """Resolve every technique's module groups against the model once."""
        model, compression_config = self.model, self.compression_config
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
        shared = e[C.SHARED_PARAMETERS]
        end = shared.get(C.TECHNIQUE_SCHEDULE_OFFSET_END) if method == C.SPARSE_PRUNING else None
        if self.training_steps >= shared[C.TECHNIQUE_SCHEDULE_OFFSET] and (end is None or end <= shared[
                C.TECHNIQUE_SCHEDULE_OFFSET] or self.training_steps <= end):
            for _, names, _ in e[C.DIFFERENT_GROUPS]:
                for n in names:
                    setattr(recursive_getattr(self.model, n), _FLAG[method], True)
            if not self.verbose[method]:
                self.verbose[method] = True
            if method == C.WEIGHT_QUANTIZATION:
                self.weight_quantization_enabled = True

    # one entry point per technique, as in the reference scheduler
    def check_weight_quantization(self):
        self._check(C.WEIGHT_QUANTIZATION)

    def check_activation_quantization(self):
        self._check(C.ACTIVATION_QUANTIZATION)

    def check_sparse_pruning(self):
        self._check(C.SPARSE_PRUNING)

    def check_head_pruning(self):
        self._check(C.HEAD_PRUNING)

    def check_row_pruning(self):
        self._check(C.ROW_PRUNING)

    def check_channel_pruning(self):
        self._check(C.CHANNEL_PRUNING)

    def check_all_modules(self):
        for check in (self.check_weight_quantization, self.check_activation_quantization, self.check_sparse_pruning,
                      self.check_head_pruning, self.check_row_pruning, self.check_channel_pruning):
            check()

    def step(self, step_zero_check=False):
        if not step_zero_check:
            self.training_steps += 1
        self.check_all_modules()
