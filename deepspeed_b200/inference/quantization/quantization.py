"""Post-init weight quantisation of a model (reference ``inference/quantization/quantization.py:23``)."""
import re

import torch

from .layers import QuantizedLinear


def _init_group_wise_weight_quantization(model, ds_config):
    """ds_config['weight_quantization']['post_init_quant'] = {pattern: {num_bits, group_size, ...}}"""
    wq = (ds_config or {}).get("weight_quantization", {}).get("post_init_quant", {})
    if not wq:
        return model
    for name, mod in list(model.named_modules()):
        for cname, child in list(mod.named_children()):
            full = f"{name}.{cname}" if name else cname
            if isinstance(child, torch.nn.Embedding):
                for pat, qc in wq.items():
                    if re.search(pat, full) and qc.get("num_bits", 8) in (4, 8) and not qc.get("fp", False):
                        from .layers import QuantizedEmbedding
                        gs = qc.get("group_size", 128)
                        if child.embedding_dim % gs == 0:
                            setattr(mod, cname, QuantizedEmbedding({"num_bits": qc.get("num_bits", 8), "group_size": gs,
                                                                    "group_dim": qc.get("group_dim", 1),
                                                                    "symmetric": False}, child))
                        break
                continue
            if not isinstance(child, torch.nn.Linear):
                continue
            for pat, qc in wq.items():
                if re.search(pat, full):
                    bits = qc.get("num_bits", 8)
                    mode = {8: "int8", 4: "int4", 6: "fp6"}[bits] if not qc.get("fp", False) else f"fp{bits}"
                    setattr(mod, cname, QuantizedLinear(child, mode, qc.get("group_size", 128)))
                    break
    return model
