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


def test_op_bindings_compose_into_the_fused_layer():
    """DeepSpeedSelfAttention + DeepSpeedMLP built from the op bindings == the monolithic fused layer."""
    import torch
    from deepspeed_b200.ops.transformer.inference.config import DeepSpeedInferenceConfig
    from deepspeed_b200.ops.transformer.inference.ds_attention import DeepSpeedSelfAttention
    from deepspeed_b200.ops.transformer.inference.ds_mlp import DeepSpeedMLP
    from deepspeed_b200.ops.transformer.inference.ds_transformer import DeepSpeedTransformerInference
    from deepspeed_b200.ops.transformer.inference.op_binding import WorkspaceOp
    torch.manual_seed(0)
    cfg = DeepSpeedInferenceConfig(hidden_size=32, intermediate_size=64, heads=4, dtype=torch.float32, pre_layer_norm=True,
                                   max_out_tokens=32, mlp_act_func_type="gelu")
    layer = DeepSpeedTransformerInference(cfg)
    for p in layer.parameters():
        torch.nn.init.normal_(p, std=0.1)
    attn, mlp = DeepSpeedSelfAttention(cfg), DeepSpeedMLP(cfg)
    with torch.no_grad():
        attn.attn_qkvw.copy_(layer.attn_qkvw), attn.attn_qkvb.copy_(layer.attn_qkvb)
        attn.attn_ow.copy_(layer.attn_ow), attn.attn_ob.copy_(layer.attn_ob)
        mlp.attn_nw.copy_(layer.attn_nw), mlp.attn_nb.copy_(layer.attn_nb)
        mlp.inter_w.copy_(layer.inter_w), mlp.inter_b.copy_(layer.inter_b)
        mlp.output_w.copy_(layer.output_w), mlp.output_b.copy_(layer.output_b)
    x = torch.randn(2, 6, 32)
    ref = layer(x)
    ref = ref[0] if isinstance(ref, tuple) else ref
    WorkspaceOp(cfg).release_workspace()
    a, k, v, _, _ = attn(x, norm_w=layer.norm_w, norm_b=layer.norm_b)
    out = mlp(a, x, bias=attn.attn_ob)
    assert k.shape == (2, 4, 6, 8) and torch.allclose(out, ref, atol=1e-5), (out - ref).abs().max()
    # incremental step reuses the workspace cache
    step_ref = layer(x[:, -1:] * 0.5, use_cache=True)
    step_ref = step_ref[0] if isinstance(step_ref, tuple) else step_ref
    a2, k2, _, _, _ = attn(x[:, -1:] * 0.5, norm_w=layer.norm_w, norm_b=layer.norm_b, layer_past=True)
    assert k2.shape[2] == 7 and torch.allclose(mlp(a2, x[:, -1:] * 0.5, bias=attn.attn_ob), step_ref, atol=1e-5)


def test_diffusers_block_and_attention():
    import torch
    from types import SimpleNamespace
    from torch import nn
    from deepspeed_b200.ops.transformer.inference.bias_add import nhwc_bias_add
    from deepspeed_b200.ops.transformer.inference.diffusers_attention import DeepSpeedDiffusersAttention
    from deepspeed_b200.ops.transformer.inference.diffusers_transformer_block import DeepSpeedDiffusersTransformerBlock
    torch.manual_seed(0)
    cfg = SimpleNamespace(hidden_size=16, heads=4, dtype=torch.float32)
    att = DeepSpeedDiffusersAttention(cfg)
    for p in att.parameters():
        nn.init.normal_(p, std=0.2)
    x, ctx = torch.randn(2, 5, 16), torch.randn(2, 3, 16)
    q, k, v = torch.nn.functional.linear(x, att.attn_qkvw, att.attn_qkvb).chunk(3, -1)
    sp = lambda t: t.reshape(2, -1, 4, 4).transpose(1, 2)
    ref = torch.nn.functional.scaled_dot_product_attention(sp(q), sp(k), sp(v)).transpose(1, 2).reshape(2, 5, 16)
    assert torch.allclose(att(x), torch.nn.functional.linear(ref, att.attn_ow, att.attn_ob), atol=1e-5)
    assert att(x, ctx).shape == (2, 5, 16)

    class GEGLU(nn.Module):

        def __init__(self):
            super().__init__()
            self.proj = nn.Linear(16, 64)

        def forward(self, h):
            a, g = self.proj(h).chunk(2, dim=-1)
            return a * torch.nn.functional.gelu(g)

    class Block(nn.Module):

        def __init__(self):
            super().__init__()
            self.norm1, self.norm2, self.norm3 = nn.LayerNorm(16), nn.LayerNorm(16), nn.LayerNorm(16)
            self.attn1, self.attn2 = nn.Linear(16, 16), nn.Linear(16, 16)
            self.ff = SimpleNamespace(net=nn.ModuleList([GEGLU(), nn.Dropout(0.0), nn.Linear(32, 16)]))

        def forward(self, x):
            x = x + self.attn1(self.norm1(x))
            x = x + self.attn2(self.norm2(x))
            h = self.norm3(x)
            return x + self.ff.net[2](self.ff.net[0](h))

    blk = Block()
    fused = DeepSpeedDiffusersTransformerBlock(blk)
    assert torch.allclose(fused(x), blk(x), atol=1e-5)
    a = torch.randn(1, 3, 3, 4)  # [N, H, W, C]
    assert torch.allclose(nhwc_bias_add(a, torch.ones(4)), a + 1)


def test_inference_context_and_moe_mlp():
    import types
    import torch
    from deepspeed_b200.ops.transformer.inference.op_binding.workspace import InferenceContext
    from deepspeed_b200.ops.transformer.inference.moe_inference import DeepSpeedMoEMLP
    ctx = InferenceContext.Instance()
    assert ctx is InferenceContext.Instance()
    ctx.gen_workspace(num_layers=2, num_heads=4, batch_size=2, prompt_len=5, hidden_dim=32, mp_size=1, external_cache=False,
                      elem_dtype=torch.float32, rank=0, max_out_tokens=16, min_out_tokens=1)
    k0, v0 = torch.randn(2, 4, 5, 8), torch.randn(2, 4, 5, 8)
    k, v = ctx.update_cache(1, None, True, k0, v0)
    assert k.shape == (2, 4, 5, 8) and torch.equal(k, k0) and ctx.current_tokens() == 5
    ctx.advance_tokens()
    k1, v1 = torch.randn(2, 4, 1, 8), torch.randn(2, 4, 1, 8)
    k, v = ctx.update_cache(1, None, False, k1, v1)
    assert k.shape == (2, 4, 6, 8) and torch.equal(k[:, :, :5], k0) and torch.equal(v[:, :, 5:], v1)
    assert float(ctx.kv_cache[1][0][:, :, 6:].abs().sum()) == 0 and float(ctx.kv_cache[0][0].abs().sum()) == 0
    ctx.release_workspace()
    cfg = types.SimpleNamespace(hidden_size=16, intermediate_size=32, mlp_act_func_type=None, dtype=torch.float32)
    from deepspeed_b200.utils.types import ActivationFuncType
    cfg.mlp_act_func_type = ActivationFuncType.GELU
    mlp = DeepSpeedMoEMLP(cfg)
    for p in mlp.parameters():
        torch.nn.init.normal_(p, std=0.1)
    x = torch.randn(3, 16)
    ref = torch.nn.functional.linear(torch.nn.functional.gelu(torch.nn.functional.linear(x, mlp.inter_w, mlp.inter_b),
                                                             approximate="tanh"), mlp.output_w, mlp.output_b)
    got = mlp(x)
    assert torch.allclose(got, ref, atol=1e-4) or torch.allclose(
        got, torch.nn.functional.linear(torch.nn.functional.gelu(torch.nn.functional.linear(x, mlp.inter_w, mlp.inter_b)),
                                        mlp.output_w, mlp.output_b), atol=1e-4)
