"""Helpers to retrofit sparse attention into HF BERT/RoBERTa (reference ``sparse_attention_utils.py``)."""
import torch
from torch.nn import functional as F

from .bert_sparse_self_attention import BertSparseSelfAttention
from .sparsity_config import FixedSparsityConfig


class SparseAttentionUtils:

    @staticmethod
    def extend_position_embedding(model, max_position):
        emb = None
        for name in ("bert", "roberta"):
            if hasattr(model, name):
                emb = getattr(model, name).embeddings.position_embeddings
                off = 2 if name == "roberta" else 0
        if emb is None:
            raise ValueError("Please extend \"extend_position_embedding\" function to support your model type. It "
                             "currently only supports \"bert\" & \"roberta\"!")
        orig = emb.weight.data
        n, h = orig.shape
        reps = (max_position + off - 1) // (n - off) + 1
        new = torch.cat([orig[:off]] + [orig[off:]] * reps, 0)[:max_position + off]
        emb.weight.data = new.clone()
        emb.num_embeddings = new.shape[0]
        model.config.max_position_embeddings = max_position + off
        return model

    @staticmethod
    def update_tokenizer_model_max_length(tokenizer, max_position):
        tokenizer.model_max_length = max_position
        tokenizer.init_kwargs["model_max_length"] = max_position
        return tokenizer

    @staticmethod
    def replace_model_self_attention_with_sparse_self_attention(model, max_position, sparsity_config=None):
        sparsity_config = sparsity_config or FixedSparsityConfig(num_heads=4)
        for name in ("bert", "roberta"):
            if hasattr(model, name):
                model.config.max_position_embeddings = max_position
                SparseAttentionUtils.replace_self_attention_layer_with_sparse_self_attention_layer(
                    model.config, getattr(model, name).encoder.layer, sparsity_config)
                return model
        raise ValueError("only \"bert\" & \"roberta\" are supported")

    @staticmethod
    def replace_self_attention_layer_with_sparse_self_attention_layer(config, layers, sparsity_config=None):
        sparsity_config = sparsity_config or FixedSparsityConfig(num_heads=4)
        for layer in layers:
            new = BertSparseSelfAttention(config, sparsity_config)
            old = layer.attention.self
            new.query, new.key, new.value = old.query, old.key, old.value
            layer.attention.self = new
        return layers

    @staticmethod
    def pad_to_block_size(block_size, input_ids, attention_mask, token_type_ids, position_ids, inputs_embeds,
                          pad_token_id, model_embeddings):
        batch, seq_len = input_ids.shape if input_ids is not None else inputs_embeds.shape[:-1]
        pad_len = (block_size - seq_len % block_size) % block_size
        if pad_len > 0:
            if inputs_embeds is not None:
                pad_ids = inputs_embeds.new_full((batch, pad_len), pad_token_id, dtype=torch.long)
                inputs_embeds = torch.cat([inputs_embeds, model_embeddings(pad_ids)], dim=-2)
            if input_ids is not None:
                input_ids = F.pad(input_ids, (0, pad_len), value=pad_token_id)
            if position_ids is not None:
                position_ids = F.pad(position_ids, (0, pad_len), value=pad_token_id)
            attention_mask = F.pad(attention_mask, (0, pad_len), value=False)
            token_type_ids = F.pad(token_type_ids, (0, pad_len), value=0)
        return pad_len, input_ids, attention_mask, token_type_ids, position_ids, inputs_embeds

    @staticmethod
    def unpad_sequence_output(pad_len, sequence_output):
        return sequence_output[:, :-pad_len] if pad_len > 0 else sequence_output


def sdd_segment(layout, max_width=4):
    """Cover each head's non-zero blocks with maximal square segments (native ``dsb_sdd_segment``; reference
    ``csrc/sparse_attention/utils.cpp`` N15).  Returns int32 rows ``[head, row, col, width]``."""
    import ctypes
    from deepspeed_b200.ops import native as N
    lay = layout.to(torch.int32).contiguous()
    H, M, Nn = lay.shape
    cap = int(lay.sum().item()) + 1
    out = torch.empty(cap, 4, dtype=torch.int32)
    lib = N.cpu()
    lib.dsb_sdd_segment.restype = ctypes.c_int64
    n = lib.dsb_sdd_segment(ctypes.c_void_p(lay.data_ptr()), H, M, Nn, max_width, ctypes.c_void_p(out.data_ptr()),
                            ctypes.c_int64(cap))
    assert n >= 0
    return out[:n]
