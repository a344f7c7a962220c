"""``deepspeed.linear.Init``: build / load a HF model with its target linears replaced by ``OptimizedLinear``
(reference ``linear/context_manager.py:23``)."""
import torch
from torch import nn

from .config import LoRAConfig, QuantizationConfig
from .optimized_linear import LoRAOptimizedLinear, OptimizedLinear


def init_lora(model):
    model.requires_grad_(False)
    for m in model.modules():
        if isinstance(m, LoRAOptimizedLinear):
            m.init_lora()


class Init:

    def __init__(self, lora_config: LoRAConfig = None, quant_config: QuantizationConfig = None):
        self._orig_linear = nn.Linear
        self.lora_config, self.quant_config = lora_config, quant_config
        self._patched = []

    def __enter__(self):
        lora, quant, orig = self.lora_config, self.quant_config, self._orig_linear

        class OptLinearWrapper:
            _orig = orig

            def __new__(cls, *args, **kwargs):
                # Like the reference wrapper (context_manager.py:50) the layer built here is bias-free unless a bias is asked
                # for EXPLICITLY; the LoRA / quantised variants have none, so an explicit bias keeps the stock layer
                # (the reference asserts instead).
                bias = kwargs.get("bias", args[2] if len(args) > 2 else False)
                dtype = kwargs.get("dtype") or torch.bfloat16
                if bias:
                    return orig(*args, **kwargs)
                return OptimizedLinear(args[0] if args else kwargs["in_features"],
                                       args[1] if len(args) > 1 else kwargs["out_features"], lora_config=lora,
                                       quantization_config=quant, dtype=dtype)

        nn.Linear = OptLinearWrapper
        try:
            import transformers

            def _post(model):
                if lora is not None and lora.delay_lora_init:
                    init_lora(model)
                return model

            for name in ("from_pretrained", "from_config"):
                fn = getattr(transformers.AutoModelForCausalLM, name)
                self._patched.append((transformers.AutoModelForCausalLM, name, fn))
                setattr(transformers.AutoModelForCausalLM, name, staticmethod(lambda *a, _f=fn, **k: _post(_f(*a, **k))))
        except Exception:
            pass
        return self

    def __exit__(self, *args):
        nn.Linear = self._orig_linear
        for obj, name, fn in self._patched:
            setattr(obj, name, fn)
        self._patched.clear()
        return False
