"""ZeRO++ on real GPUs: int8 weight all-gather (qwZ), hierarchical secondary partition (hpZ) and int4 quantised gradient
reduction (qgZ) run the device quantisation kernels of ``csrc/cuda/quant.cu`` between ranks; the loss must track the plain
ZeRO-3 run within quantisation noise.  Also: the NVMe optimizer tier against the box's local disk."""
import os

import pytest
import torch

from tests.common import run_distributed

pytestmark = pytest.mark.gpu


def _need(n):
    if not torch.cuda.is_available() or torch.cuda.device_count() < n:
        pytest.skip(f"needs {n} GPUs")


def _train(zero_extra, steps=8):
    import deepspeed_b200 as ds
    from deepspeed_b200.models.llama import LlamaForCausalLM, llama_config
    r, w = ds.comm.get_rank(), ds.comm.get_world_size()
    cfg = llama_config("tiny", hidden_size=256, intermediate_size=512, num_attention_heads=4, num_key_value_heads=2,
                       vocab_size=1024, num_hidden_layers=2)
    torch.manual_seed(0)
    with torch.device("cuda"):
        model = LlamaForCausalLM(cfg).to(torch.bfloat16)
    zero = {"stage": 3, "stage3_param_persistence_threshold": 0}
    zero.update(zero_extra)
    conf = {"train_micro_batch_size_per_gpu": 2, "bf16": {"enabled": True}, "zero_optimization": zero,
            "optimizer": {"type": "AdamW", "params": {"lr": 2e-3}}}
    eng, *_ = ds.initialize(model=model, config=conf)
    g = torch.Generator().manual_seed(7)
    ids_all = torch.randint(0, cfg.vocab_size, (2 * w, 64), generator=g)  # one fixed batch: the loss must go down
    ids = ids_all[r * 2:(r + 1) * 2].cuda()
    losses = []
    for _ in range(steps):
        loss = eng(ids, labels=ids)
        eng.backward(loss)
        eng.step()
        losses.append(loss.item())
    eng.destroy()
    return losses


def _zeropp_worker():
    from deepspeed_b200.ops import native as N
    base = _train({})
    assert base[-1] < base[0] - 0.2, base
    for extra, tol in (({"zero_quantized_weights": True}, 0.25), ({"zero_quantized_gradients": True}, 0.35),
                       ({"zero_hpz_partition_size": 2}, 0.05),
                       ({"zero_quantized_weights": True, "zero_quantized_gradients": True, "zero_hpz_partition_size": 2}, 0.45)):
        n0 = N.launch_count
        got = _train(extra)
        assert N.launch_count > n0, "native kernels must have run"
        assert got[-1] < got[0] - 0.15, (extra, got)
        assert abs(got[-1] - base[-1]) < tol, (extra, got[-1], base[-1])


@pytest.mark.parametrize("world", [2, 4])
def test_zeropp_tracks_plain_zero3(world):
    _need(world)
    run_distributed(_zeropp_worker, world, backend="nccl", timeout=600)


def _nvme_worker(path):
    import deepspeed_b200 as ds
    from deepspeed_b200.utils import safe_get_full_fp32_param
    finals = {}
    for mode in ("cpu", "nvme"):
        torch.manual_seed(0)
        model = torch.nn.Sequential(torch.nn.Linear(256, 512), torch.nn.GELU(), torch.nn.Linear(512, 256)).cuda().bfloat16()
        off = {"device": "cpu", "pin_memory": True} if mode == "cpu" else {"device": "nvme", "nvme_path": path,
                                                                             "pin_memory": True, "b200_swap_window": 65536}
        conf = {"train_micro_batch_size_per_gpu": 4, "bf16": {"enabled": True},
                "optimizer": {"type": "AdamW", "params": {"lr": 1e-3}},
                "zero_optimization": {"stage": 3, "offload_optimizer": off}}
        eng, *_ = ds.initialize(model=model, config=conf)
        g = torch.Generator().manual_seed(3)
        for _ in range(4):
            x = torch.randn(4, 256, generator=g).cuda().bfloat16()
            loss = eng(x).float().pow(2).mean()
            eng.backward(loss)
            eng.step()
        finals[mode] = [safe_get_full_fp32_param(p).cpu().clone() for p in model.parameters()]
        eng.destroy()
    assert any(f.endswith(".swp") for _, _, fs in os.walk(path) for f in fs), "no swap files were written"
    for a, b in zip(finals["cpu"], finals["nvme"]):
        torch.testing.assert_close(a, b, atol=1e-6, rtol=1e-5)


def test_nvme_optimizer_tier_on_local_disk(tmp_path):
    """ZeRO-Infinity optimizer swap (O_DIRECT libaio files) on the GPU box's own disk == the pinned-host tier."""
    _need(1)
    run_distributed(_nvme_worker, 1, args=(str(tmp_path / "swap"), ), backend="nccl", timeout=300)
