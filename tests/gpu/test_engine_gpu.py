"""GPU tier: engine + flagship model end to end on one B200."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_native_library_is_loaded():
    from deepspeed_b200.ops import native
    lib = native.cuda()
    assert lib is not None
    maps = open("/proc/self/maps").read()
    assert "libdsb200_cuda.so" in maps


def test_llama_bf16_matches_hf_on_gpu():
    from transformers import LlamaConfig as HFC, LlamaForCausalLM as HFL
    from deepspeed_b200.models.llama import LlamaForCausalLM, llama_config
    torch.manual_seed(0)
    cfg = llama_config("tiny", hidden_size=256, intermediate_size=512, num_attention_heads=4, num_key_value_heads=2,
                       vocab_size=1024, num_hidden_layers=2)
    hf = HFL(HFC(vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
                 num_hidden_layers=cfg.num_hidden_layers, num_attention_heads=cfg.num_attention_heads,
                 num_key_value_heads=cfg.num_key_value_heads, max_position_embeddings=cfg.max_position_embeddings,
                 rms_norm_eps=cfg.rms_norm_eps, tie_word_embeddings=False)).cuda()
    m = LlamaForCausalLM(cfg).cuda()
    m.load_state_dict(LlamaForCausalLM.convert_hf_state_dict(hf.state_dict(), cfg))
    ids = torch.randint(0, cfg.vocab_size, (2, 128), device="cuda")
    ref = hf(input_ids=ids, labels=ids).loss  # fp32 reference
    mb = m.to(torch.bfloat16)
    loss = mb(ids, labels=ids)
    assert abs(loss.item() - ref.item()) < 5e-2, (loss.item(), ref.item())
    loss.backward()
    ref.backward()
    g = mb.model.layers[1].mlp.down_proj.weight.grad.float()
    gr = hf.model.layers[1].mlp.down_proj.weight.grad
    cos = torch.nn.functional.cosine_similarity(g.flatten(), gr.flatten(), dim=0)
    assert cos > 0.99, cos


@pytest.mark.parametrize("stage", [0, 2, 3])
def test_engine_trains_tiny_llama(stage):
    import deepspeed_b200 as ds
    from deepspeed_b200.models.llama import LlamaForCausalLM, llama_config
    from deepspeed_b200.ops import native
    torch.manual_seed(0)
    cfg = llama_config("tiny", hidden_size=256, intermediate_size=512, num_attention_heads=4, num_key_value_heads=2,
                       vocab_size=1024, num_hidden_layers=2, checkpoint_layers=1)
    with torch.device("cuda"):
        model = LlamaForCausalLM(cfg).to(torch.bfloat16)
    eng, _, _, _ = ds.initialize(model=model, config={
        "train_micro_batch_size_per_gpu": 2, "bf16": {"enabled": True},
        "optimizer": {"type": "AdamW", "params": {"lr": 2e-3}},
        "zero_optimization": {"stage": stage}, "gradient_clipping": 1.0 if stage == 2 else 0.0})
    ids = torch.randint(0, cfg.vocab_size, (2, 128), device="cuda")
    before = native.launch_count
    losses = []
    for _ in range(8):
        loss = eng(ids, labels=ids)
        eng.backward(loss)
        eng.step()
        losses.append(loss.item())
    assert losses[-1] < losses[0] - 0.5, losses
    assert native.launch_count - before > 50
    if stage == 2:
        assert eng.get_global_grad_norm() is not None and eng.get_global_grad_norm() > 0


def test_checkpoint_roundtrip_gpu(tmp_path):
    import deepspeed_b200 as ds
    from deepspeed_b200.models.llama import LlamaForCausalLM, llama_config
    cfg = llama_config("tiny")
    conf = {"train_micro_batch_size_per_gpu": 2, "bf16": {"enabled": True},
            "optimizer": {"type": "AdamW", "params": {"lr": 1e-3}}, "zero_optimization": {"stage": 3}}
    torch.manual_seed(0)
    with torch.device("cuda"):
        m1 = LlamaForCausalLM(cfg).to(torch.bfloat16)
    e1, _, _, _ = ds.initialize(model=m1, config=conf)
    ids = torch.randint(0, cfg.vocab_size, (2, 64), device="cuda")
    for _ in range(2):
        loss = e1(ids, labels=ids)
        e1.backward(loss)
        e1.step()
    e1.save_checkpoint(str(tmp_path), tag="t1", client_state={"foo": 7})
    with torch.device("cuda"):
        m2 = LlamaForCausalLM(cfg).to(torch.bfloat16)
    e2, _, _, _ = ds.initialize(model=m2, config=conf)
    path, client = e2.load_checkpoint(str(tmp_path))
    assert path is not None and client["foo"] == 7 and e2.global_steps == 2
    l1 = e1(ids, labels=ids)
    l2 = e2(ids, labels=ids)
    assert abs(l1.item() - l2.item()) < 1e-3
    e1.backward(l1); e1.step()
    e2.backward(l2); e2.step()
    l1b, l2b = e1(ids, labels=ids).item(), e2(ids, labels=ids).item()
    assert abs(l1b - l2b) < 2e-3
