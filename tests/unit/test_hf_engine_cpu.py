"""The engine on an UNMODIFIED Hugging Face ``LlamaForCausalLM`` (the module the reference arm of bench.py trains):
ZeRO-2/3 over gloo world_size=2 with clipping must track a plain torch AdamW run."""
import copy

import pytest
import torch

from tests.common import run_distributed

def _worker(stage):
    import deepspeed_b200 as ds
    from deepspeed_b200.utils import safe_get_full_fp32_param
    from transformers import LlamaConfig, LlamaForCausalLM
    torch.manual_seed(0)
    cfg = LlamaConfig(vocab_size=256, hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
                      num_key_value_heads=2, max_position_embeddings=64, use_cache=False, tie_word_embeddings=False)
    m = LlamaForCausalLM(cfg); m.train()
    ref = copy.deepcopy(m)
    conf = {"train_micro_batch_size_per_gpu": 2, "optimizer": {"type": "AdamW", "params": {"lr": 1e-3, "weight_decay": 0.1}},
            "gradient_clipping": 1.0,
            "zero_optimization": {"stage": stage, "stage3_param_persistence_threshold": 0}}
    eng, *_ = ds.initialize(model=m, config=conf)
    r, w = ds.comm.get_rank(), ds.comm.get_world_size()
    ropt = torch.optim.AdamW(ref.parameters(), lr=1e-3, weight_decay=0.1)
    g = torch.Generator().manual_seed(1)
    for _ in range(3):
        ids = torch.randint(0, 256, (2 * w, 16), generator=g)
        loss = eng(input_ids=ids[r*2:(r+1)*2], labels=ids[r*2:(r+1)*2]).loss
        eng.backward(loss); eng.step()
        rl = sum(ref(input_ids=ids[k*2:(k+1)*2], labels=ids[k*2:(k+1)*2]).loss for k in range(w)) / w
        rl.backward(); torch.nn.utils.clip_grad_norm_(ref.parameters(), 1.0); ropt.step(); ropt.zero_grad()
    worst = max((safe_get_full_fp32_param(p).cpu() - q).abs().max().item() for p, q in zip(m.parameters(), ref.parameters()))
    assert worst < 5e-5, worst


@pytest.mark.parametrize("stage", [2, 3])
def test_engine_trains_hf_llama(stage):
    run_distributed(_worker, 2, (stage, ))
