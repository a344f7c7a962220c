import copy

import torch
import torch.nn.functional as F

from deepspeed_b200.ops.transformer import (DeepSpeedInferenceConfig, DeepSpeedTransformerConfig, DeepSpeedTransformerInference,
                                            DeepSpeedTransformerLayer)


def _ref_layer(l, x, mask, pre_ln):
    c = l.config
    B, S, H = x.shape
    nh, d = c.heads, H // c.heads
    ln = lambda t, w, b: F.layer_norm(t, (H, ), w, b, c.layer_norm_eps)
    a_in = ln(x, l.attn_nw, l.attn_nb) if pre_ln else x
    qkv = F.linear(a_in, l.attn_qkvw, l.attn_qkvb).view(B, S, 3, nh, d).permute(2, 0, 3, 1, 4)
    sc = qkv[0] @ qkv[1].transpose(-1, -2) / d**0.5 + (mask if mask is not None else 0)
    ctx = (torch.softmax(sc, -1) @ qkv[2]).transpose(1, 2).reshape(B, S, H)
    x1 = x + F.linear(ctx, l.attn_ow, l.attn_ob)
    if pre_ln:
        f_in = ln(x1, l.norm_w, l.norm_b)
    else:
        x1 = ln(x1, l.attn_nw, l.attn_nb)
        f_in = x1
    out = x1 + F.linear(F.gelu(F.linear(f_in, l.inter_w, l.inter_b)), l.output_w, l.output_b)
    return out if pre_ln else ln(out, l.norm_w, l.norm_b)


def test_training_layer_matches_reference_math():
    for pre_ln in (True, False):
        torch.manual_seed(0)
        cfg = DeepSpeedTransformerConfig(batch_size=2, hidden_size=32, intermediate_size=64, heads=4, attn_dropout_ratio=0.0,
                                         hidden_dropout_ratio=0.0, num_hidden_layers=2, initializer_range=0.02,
                                         pre_layer_norm=pre_ln, gelu_checkpoint=True, attn_dropout_checkpoint=True)
        l = DeepSpeedTransformerLayer(cfg)
        for b in (l.attn_qkvb, l.attn_ob, l.inter_b, l.output_b):
            b.data.normal_(0, 0.1)
        x = torch.randn(2, 6, 32, requires_grad=True)
        mask = torch.zeros(2, 1, 1, 6)
        mask[:, :, :, -2:] = -10000.0
        out = l(x, mask)
        ref = _ref_layer(l, x, mask, pre_ln)
        torch.testing.assert_close(out, ref, atol=1e-5, rtol=1e-4)
        g = torch.autograd.grad(out.sum(), [x, l.attn_qkvw, l.inter_b])
        gr = torch.autograd.grad(ref.sum(), [x, l.attn_qkvw, l.inter_b])
        for a, b in zip(g, gr):
            torch.testing.assert_close(a, b, atol=1e-4, rtol=1e-3)
    # dropout path runs and keeps the expected scale
    cfg = DeepSpeedTransformerConfig(batch_size=2, hidden_size=32, heads=4, attn_dropout_ratio=0.1, hidden_dropout_ratio=0.1,
                                     num_hidden_layers=2, initializer_range=0.02, seed=7)
    l = DeepSpeedTransformerLayer(cfg).train()
    y = l(torch.randn(2, 6, 32))
    assert torch.isfinite(y).all()


def test_inference_layer_kv_cache_decode_matches_full():
    torch.manual_seed(0)
    cfg = DeepSpeedInferenceConfig(hidden_size=32, intermediate_size=64, heads=4, num_hidden_layers=1, dtype=torch.float32,
                                   rotary_dim=8, max_out_tokens=32, mlp_act_func_type="gelu", num_kv=2)
    l = DeepSpeedTransformerInference(cfg)
    for p in l.parameters():
        p.data.normal_(0, 0.1)
    l.norm_w.data.fill_(1.0), l.attn_nw.data.fill_(1.0)
    x = torch.randn(2, 7, 32)
    full, _ = l(x)
    l.reset_cache()
    a, _ = l(x[:, :5], use_cache=True)
    outs = [a]
    for t in range(5, 7):
        o, _ = l(x[:, t:t + 1], use_cache=True)
        outs.append(o)
    torch.testing.assert_close(torch.cat(outs, 1), full, atol=1e-5, rtol=1e-4)
