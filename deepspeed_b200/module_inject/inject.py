"""Training-time layer replacement (reference ``module_inject/inject.py:module_inject``): swap every
``layer_obj`` (a HF BERT layer) for the fused *training* transformer layer and copy its weights."""
import torch

from deepspeed_b200.ops.transformer import DeepSpeedTransformerConfig, DeepSpeedTransformerLayer


def module_inject(layer_obj, model, config, micro_batch_size, max_seq_length, seed, preln, fp16=True):
    for name, child in list(model.named_children()):
        if isinstance(child, layer_obj):
            cfg = DeepSpeedTransformerConfig(batch_size=micro_batch_size, hidden_size=config.hidden_size,
                                             intermediate_size=config.intermediate_size, heads=config.num_attention_heads,
                                             attn_dropout_ratio=config.attention_probs_dropout_prob,
                                             hidden_dropout_ratio=config.hidden_dropout_prob,
                                             num_hidden_layers=config.num_hidden_layers, initializer_range=config.initializer_range,
                                             layer_norm_eps=getattr(config, "layer_norm_eps", 1e-12), seed=seed, fp16=fp16,
                                             pre_layer_norm=preln)
            new = DeepSpeedTransformerLayer(cfg)
            a = child.attention
            with torch.no_grad():
                new.attn_qkvw.copy_(torch.cat([a.self.query.weight, a.self.key.weight, a.self.value.weight], 0))
                new.attn_qkvb.copy_(torch.cat([a.self.query.bias, a.self.key.bias, a.self.value.bias], 0))
                new.attn_ow.copy_(a.output.dense.weight)
                new.attn_ob.copy_(a.output.dense.bias)
                ln_attn = child.PostAttentionLayerNorm if preln and hasattr(child, "PostAttentionLayerNorm") else a.output.LayerNorm
                new.attn_nw.copy_(ln_attn.weight)
                new.attn_nb.copy_(ln_attn.bias)
                inter = child.intermediate.dense_act if hasattr(child.intermediate, "dense_act") else child.intermediate.dense
                new.inter_w.copy_(inter.weight)
                new.inter_b.copy_(inter.bias)
                new.output_w.copy_(child.output.dense.weight)
                new.output_b.copy_(child.output.dense.bias)
                ln_out = child.PreAttentionLayerNorm if preln and hasattr(child, "PreAttentionLayerNorm") else child.output.LayerNorm
                new.norm_w.copy_(ln_out.weight)
                new.norm_b.copy_(ln_out.bias)
            setattr(model, name, new)
        else:
            module_inject(layer_obj, child, config, micro_batch_size, max_seq_length, seed, preln, fp16)
    return model
