"""v1 inference engine (kernel injection, generate), AutoTP sharding (ws=2 gloo), ragged TP=2, Domino."""
import pytest
import torch

from tests.common import run_distributed

transformers = pytest.importorskip("transformers")


def _tiny_llama():
    from transformers import AutoConfig, AutoModelForCausalLM
    cfg = AutoConfig.for_model("llama", vocab_size=128, hidden_size=64, num_hidden_layers=2, num_attention_heads=4,
                               num_key_value_heads=2, intermediate_size=96, max_position_embeddings=256, eos_token_id=None,
                               bos_token_id=None, pad_token_id=None)
    torch.manual_seed(0)
    return AutoModelForCausalLM.from_config(cfg).float().eval()


def test_init_inference_kernel_inject_forward_generate():
    import deepspeed_b200 as ds
    m = _tiny_llama()
    eng = ds.init_inference(m, dtype=torch.float32, replace_with_kernel_inject=True, max_out_tokens=64)
    ids = torch.randint(0, 128, (2, 9))
    out = eng(input_ids=ids)
    with torch.no_grad():
        ref = m(ids).logits
    torch.testing.assert_close(out.logits, ref, atol=2e-4, rtol=1e-3)
    gen = eng.generate(ids, max_new_tokens=6)
    with torch.no_grad():
        cur = ids
        for _ in range(6):
            cur = torch.cat([cur, m(cur).logits[:, -1].argmax(-1, keepdim=True)], 1)
    assert torch.equal(gen, cur)


def test_init_inference_plain():
    import deepspeed_b200 as ds
    m = _tiny_llama()
    eng = ds.init_inference(m, dtype=torch.float32)
    ids = torch.randint(0, 128, (1, 5))
    torch.testing.assert_close(eng(ids).logits, m(ids).logits)


def _autotp_worker():
    import copy
    import torch.distributed as td
    import deepspeed_b200 as ds
    from deepspeed_b200.module_inject import LinearAllreduce, LinearLayer
    m = _tiny_llama()
    ref = copy.deepcopy(m)
    ids = torch.randint(0, 128, (2, 7), generator=torch.Generator().manual_seed(3))
    sharded = ds.tp_model_init(m, tp_size=2, dtype=torch.float32)
    kinds = [type(x) for x in sharded.modules()]
    assert LinearAllreduce in kinds and LinearLayer in kinds
    q = sharded.model.layers[0].self_attn.q_proj
    assert q.weight.shape == (32, 64)
    with torch.no_grad():
        out, want = sharded(ids).logits, ref(ids).logits
    torch.testing.assert_close(out, want, atol=2e-4, rtol=1e-3)
    # training step: sharded grads equal the matching slices of the reference grads
    sharded.train(), ref.train()
    sharded(ids).logits.float().pow(2).mean().backward()
    ref(ids).logits.float().pow(2).mean().backward()
    r = td.get_rank()
    g = ref.model.layers[0].self_attn.q_proj.weight.grad[r * 32:(r + 1) * 32]
    torch.testing.assert_close(q.weight.grad, g, atol=1e-5, rtol=1e-3)
    go = ref.model.layers[1].mlp.down_proj.weight.grad[:, r * 48:(r + 1) * 48]
    torch.testing.assert_close(sharded.model.layers[1].mlp.down_proj.weight.grad, go, atol=1e-5, rtol=1e-3)


def test_autotp_ws2():
    run_distributed(_autotp_worker, 2)


def _ragged_tp_worker():
    from deepspeed_b200.inference.v2 import build_hf_engine
    m = _tiny_llama()
    cfg = {"tensor_parallel": {"tp_size": 2}, "state_manager": {"max_context": 128, "max_ragged_batch_size": 64,
           "max_ragged_sequence_count": 4, "memory_config": {"mode": "allocate", "size": 8}}}
    e = build_hf_engine(m, cfg, dtype=torch.float32, device="cpu")
    ids = torch.randint(0, 128, (9, ), generator=torch.Generator().manual_seed(5))
    got = e.put([0], [ids])[0]
    with torch.no_grad():
        ref = m(ids[None]).logits[0, -1]
    torch.testing.assert_close(got, ref, atol=2e-4, rtol=1e-3)
    nxt = got.argmax().reshape(1)
    got2 = e.put([0], [nxt])[0]
    with torch.no_grad():
        ref2 = m(torch.cat([ids, nxt])[None]).logits[0, -1]
    torch.testing.assert_close(got2, ref2, atol=2e-4, rtol=1e-3)


def test_ragged_engine_tp2():
    run_distributed(_ragged_tp_worker, 2)


def _domino_worker():
    import torch.distributed as td
    from deepspeed_b200.runtime.domino import DominoTransformerLayer
    torch.manual_seed(0)
    full = DominoTransformerLayer(32, 4, 64, tp_group=None)
    r, w = td.get_rank(), td.get_world_size()
    tp = DominoTransformerLayer(32, 4, 64, tp_group=td.group.WORLD)
    with torch.no_grad():
        hd = 8
        hl = 4 // w
        qkv_w = full.self_attention.qkv.weight.view(3, 4, hd, 32)[:, r * hl:(r + 1) * hl].reshape(-1, 32)
        qkv_b = full.self_attention.qkv.bias.view(3, 4, hd)[:, r * hl:(r + 1) * hl].reshape(-1)
        tp.self_attention.qkv.weight.copy_(qkv_w), tp.self_attention.qkv.bias.copy_(qkv_b)
        tp.self_attention.dense.weight.copy_(full.self_attention.dense.weight[:, r * hl * hd:(r + 1) * hl * hd])
        f = 64 // w
        tp.mlp.fc1.weight.copy_(full.mlp.fc1.weight[r * f:(r + 1) * f]), tp.mlp.fc1.bias.copy_(full.mlp.fc1.bias[r * f:(r + 1) * f])
        tp.mlp.fc2.weight.copy_(full.mlp.fc2.weight[:, r * f:(r + 1) * f])
        for a, b in ((tp.input_layernorm, full.input_layernorm), (tp.post_attention_layernorm, full.post_attention_layernorm)):
            a.load_state_dict(b.state_dict())
    x = torch.randn(4, 6, 32, generator=torch.Generator().manual_seed(1))
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    ya, yb = tp(xa), full(xb)
    torch.testing.assert_close(ya, yb, atol=1e-5, rtol=1e-4)
    ya.sum().backward(), yb.sum().backward()
    torch.testing.assert_close(xa.grad, xb.grad, atol=1e-5, rtol=1e-4)
    # every async backward all-reduce was drained by its NoOper node; weight grads match the dense slice
    from deepspeed_b200.runtime.domino import transformer as D
    assert not D.handle_dic
    torch.testing.assert_close(tp.mlp.fc1.weight.grad, full.mlp.fc1.weight.grad[r * f:(r + 1) * f], atol=1e-5, rtol=1e-4)
    torch.testing.assert_close(tp.input_layernorm.weight.grad, full.input_layernorm.weight.grad, atol=1e-5, rtol=1e-4)

    class _Cfg:
        attention_dropout, kv_channels, num_attention_heads = 0.0, 8, 4

    class _Mpu:
        get_tensor_model_parallel_world_size = staticmethod(lambda: w)
        get_tensor_model_parallel_group = staticmethod(lambda: td.group.WORLD)

    core = D.CoreAttention(_Cfg, 1, _Mpu)
    q = torch.randn(2, 4 // w, 5, 8)
    assert core(q, q, q, None).shape == (5, 2, 32 // w)
    assert D.AttnMaskType.causal.value == 2 and D.LayerType.encoder.value == 1 and D.ModelType.encoder_or_decoder.value == 1
    t = torch.ones(3, requires_grad=True)
    y = D.copy_to_tensor_model_parallel_region_a(_Mpu, D.no_oper(t * 1.0, D.handle_dic, "k"), D.handle_dic, "k")
    y.sum().backward()
    assert torch.equal(t.grad, torch.full((3, ), float(w))) and not D.handle_dic


def test_domino_tp2_matches_dense():
    run_distributed(_domino_worker, 2)


def _tp_variants():
    import torch
    import torch.distributed as td
    from torch import nn
    from deepspeed_b200.module_inject import layers as L
    r, w = td.get_rank(), td.get_world_size()
    g = td.group.WORLD
    torch.manual_seed(0)
    x = torch.randn(3, 16)
    # fused [q|k|v] (glm layout), gate/up pack, Conv1D column + row
    fused = nn.Linear(16, 48)
    lay = L.fused_LinearLayer(fused, g, fused_type="glmtype")
    full = fused(x)
    q, k, v = full.chunk(3, -1)
    want = torch.cat([t.chunk(w, -1)[r] for t in (q, k, v)], -1)
    assert torch.allclose(lay(x), want, atol=1e-6)
    gu = nn.Linear(16, 32, bias=False)
    lay = L.GateUpPack_LinearLayer(gu, g)
    a, b = gu(x).chunk(2, -1)
    assert torch.allclose(lay(x), torch.cat([a.chunk(w, -1)[r], b.chunk(w, -1)[r]], -1), atol=1e-6)

    class Conv1D(nn.Module):

        def __init__(self, nin, nout):
            super().__init__()
            self.weight, self.bias = nn.Parameter(torch.randn(nin, nout) * 0.1), nn.Parameter(torch.randn(nout) * 0.1)

        def forward(self, t):
            return t @ self.weight + self.bias

    c1, c2 = Conv1D(16, 32), Conv1D(32, 16)
    col, row = L.conv_LinearLayer(c1, g), L.Conv_LinearALlreduce(c2, g)
    assert torch.allclose(row(col(x)), c2(c1(x)), atol=1e-5)
    # output-channel sharded conv feeding an input-channel sharded conv == the two full convs
    torch.manual_seed(1)
    conv_a, conv_b = nn.Conv2d(3, 8, 3, padding=1), nn.Conv2d(8, 4, 1)
    img = torch.randn(2, 3, 6, 6)
    ref = conv_b(conv_a(img))
    import copy
    oc = L.TensorParallelOcShardConv2d(copy.deepcopy(conv_a), r, w)
    ic = L.TensorParallelIcShardConv2d(copy.deepcopy(conv_b), r, w, g)
    assert torch.allclose(ic(oc(img)), ref, atol=1e-5)
    assert L.get_auto_tp_mode().value == "INFERENCE" and not L.is_autotp_training_mode()
    n = L.RMSNormalize(dim=16, dtype=torch.float32)
    assert torch.allclose(n(x), x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + 1e-5), atol=1e-5)
    emb = L.OPTEmbedding(weight=nn.Parameter(torch.arange(20.).reshape(10, 2)))
    out = emb(torch.tensor([[0, 1, 1, 1]]))
    assert out.shape == (1, 4, 2) and out[0, 1, 0] == 4.0  # first real token -> position 0 -> row 2


def test_tp_layer_variants():
    run_distributed(_tp_variants, 2)
