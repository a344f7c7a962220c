"""``QuantizationContext``: construct a model under ZeRO-3 ``Init`` with its weights quantized as they are created, so
the full-precision model never exists in memory (reference ``inference/quantization/quantization_context.py``)."""
import json

from deepspeed_b200.runtime.zero.partition_parameters import Init, is_zero_param


class QuantizationContext(Init):

    def __init__(self, config_dict_or_path, param_swapper=None) -> None:
        super().__init__(config_dict_or_path=config_dict_or_path, param_swapper=param_swapper)
        cfg = config_dict_or_path
        if isinstance(cfg, str):
            with open(cfg) as f:
                cfg = json.load(f)
        self.weight_quantization_config = (cfg or {}).get("weight_quantization", {}).get("post_init_quant", {})

    def _quant_conf(self, qualified_name):
        for key, qc in self.weight_quantization_config.items():
            if key in qualified_name:
                return {"num_bits": qc.get("num_bits", 8), "group_size": qc.get("group_size", 64),
                        "group_dim": qc.get("group_dim", 1), "symmetric": qc.get("symmetric", False)}
        return None

    def _shard_module(self, m):
        """Quantize matching 2-D weights (keys match against ``<ClassName>.<param path>``) and only then shard them."""
        from .utils import _quantize_param
        for name, p in m.named_parameters(recurse=True):
            if is_zero_param(p) or getattr(p, "weight_quantized", False):
                continue
            conf = self._quant_conf(f"{type(m).__name__}.{name}")
            if conf is not None and p.dim() >= 2 and p.shape[conf["group_dim"]] % conf["group_size"] == 0:
                p.quant_full_shape = tuple(p.shape)
                _quantize_param(p, conf)
        super()._shard_module(m)
