"""Activation checkpointing, comm facade + comms logger, monitors, dataloaders, timers (host tier)."""
import os

import pytest
import torch
from torch import nn

from tests.common import run_distributed


def test_activation_checkpointing_matches_plain_and_rng():
    from deepspeed_b200.runtime.activation_checkpointing import checkpointing as ck
    ck.reset()
    ck.configure(None, partition_activations=False, checkpoint_in_cpu=False)
    torch.manual_seed(0)
    lin1, lin2 = nn.Linear(16, 32), nn.Linear(32, 16)
    drop = nn.Dropout(0.3)

    def block(x, scale):
        return lin2(drop(torch.relu(lin1(x)))) * scale

    x = torch.randn(4, 16, requires_grad=True)
    torch.manual_seed(5)
    y = ck.checkpoint(block, x, 2.0)
    y.sum().backward()
    g_ck, gw_ck = x.grad.clone(), lin1.weight.grad.clone()
    x.grad = None
    lin1.weight.grad = None
    lin2.weight.grad = None
    torch.manual_seed(5)
    y2 = block(x, 2.0)
    y2.sum().backward()
    torch.testing.assert_close(y, y2)            # same dropout mask: RNG state is captured and replayed
    torch.testing.assert_close(g_ck, x.grad)
    torch.testing.assert_close(gw_ck, lin1.weight.grad)
    # CPU checkpointing + non-tensor / multiple outputs
    ck.configure(None, checkpoint_in_cpu=True)

    def two(x):
        return x * 2, x.sum()

    a, b = ck.checkpoint(two, x)
    (a.sum() + b).backward()
    assert ck.is_configured()
    ck.reset()


def _comm_worker():
    import deepspeed_b200.comm as dist
    dist.configure(enabled=True, prof_all=True, verbose=False)
    r, w = dist.get_rank(), dist.get_world_size()
    t = torch.full((8, ), float(r + 1))
    dist.all_reduce(t)
    assert torch.all(t == sum(range(1, w + 1)))
    out = torch.empty(8 * w)
    dist.all_gather_into_tensor(out, torch.full((8, ), float(r)))
    assert torch.equal(out, torch.arange(w).float().repeat_interleave(8))
    rs = torch.empty(4)
    dist.reduce_scatter_tensor(rs, torch.arange(4 * w).float())
    assert torch.equal(rs, torch.arange(4 * w).float()[r * 4:(r + 1) * 4] * w)
    a2a = torch.empty(2 * w)
    dist.all_to_all_single(a2a, torch.arange(2 * w).float() + 100 * r)
    assert a2a.view(w, 2)[:, 0].tolist() == [100 * s + 2 * r for s in range(w)]
    b = torch.tensor([float(r)])
    dist.broadcast(b, src=1)
    assert b.item() == 1.0
    objs = [None] * w
    dist.all_gather_object(objs, {"rank": r})
    assert [o["rank"] for o in objs] == list(range(w))
    dist.barrier()
    dist.monitored_barrier()
    lg = dist.comms_logger
    assert lg is not None and any("all_reduce" in k for k in lg.comms_dict), list(lg.comms_dict)
    dist.log_summary()
    os.environ["DSB200_COMM_ALL_REDUCE_OFF"] = "1"
    z = torch.ones(2)
    dist.all_reduce(z)                      # kill switch: op becomes a no-op
    assert torch.all(z == 1)
    del os.environ["DSB200_COMM_ALL_REDUCE_OFF"]


def test_comm_facade_and_logger_ws2():
    run_distributed(_comm_worker, 2)


def test_csv_monitor_and_master(tmp_path):
    from deepspeed_b200.monitor.monitor import MonitorMaster
    from deepspeed_b200.runtime.config import DeepSpeedConfig
    cfg = DeepSpeedConfig({"train_batch_size": 1, "csv_monitor": {"enabled": True, "output_path": str(tmp_path),
                                                                  "job_name": "job"}})
    mon = MonitorMaster(cfg.monitor_config)
    assert mon.enabled
    mon.write_events([("Train/Samples/lr", 0.1, 1), ("Train/Samples/train_loss", 2.5, 1)])
    mon.write_events([("Train/Samples/lr", 0.05, 2)])
    txt = (tmp_path / "job" / "Train_Samples_lr.csv").read_text().strip().splitlines()
    assert txt[0] == "step,lr" and txt[1:] == ["1,0.1", "2,0.05"]


def test_dataloaders():
    from deepspeed_b200.runtime.dataloader import DeepSpeedDataLoader, RepeatingLoader
    ds = torch.utils.data.TensorDataset(torch.arange(10).float())
    dl = DeepSpeedDataLoader(ds, batch_size=4, pin_memory=False, local_rank=0, tput_timer=None,
                             data_parallel_world_size=2, data_parallel_rank=1, dataloader_drop_last=True)
    batches = list(dl)
    assert len(batches) == 1 and batches[0][0].numel() == 4          # 5 samples on this rank, drop_last
    rep = RepeatingLoader([1, 2, 3])
    it = iter(rep)
    assert [next(it) for _ in range(7)] == [1, 2, 3, 1, 2, 3, 1]


def test_timers_and_throughput():
    from deepspeed_b200.utils.timer import SynchronizedWallClockTimer, ThroughputTimer
    t = SynchronizedWallClockTimer()
    t("fwd").start()
    sum(range(10000))
    t("fwd").stop()
    assert t("fwd").elapsed(reset=False) >= 0 and "fwd" in t.get_timers()
    means = t.get_mean(["fwd"], reset=True)
    assert "fwd" in means
    tp = ThroughputTimer(batch_size=8, start_step=1)
    for _ in range(4):
        tp.start()
        tp.stop(global_step=True)
    assert tp.global_step_count == 4 and tp.avg_samples_per_sec() > 0


def test_contiguous_memory_allocator_defragments():
    import torch
    from deepspeed_b200.runtime.zero.contiguous_memory_allocator import ContiguousMemoryAllocator
    a = ContiguousMemoryAllocator(64, torch.float32, "cpu")
    ts = [a.allocate_tensor(16) for _ in range(4)]
    for i, t in enumerate(ts):
        t.fill_(float(i + 1))
    p = torch.nn.Parameter(torch.empty(0))
    a.assign_to_param(ts[3], p, 12, (3, 4))
    assert p.shape == (3, 4) and float(p.sum()) == 48.0
    a.release_tensor(ts[0])
    a.release_tensor(ts[2])
    assert a.total_free == 32 and a.largest_contiguous == 16
    big = a.allocate_tensor(32)  # needs compaction: two 16-element holes
    assert big.numel() == 32 and a.total_free == 0 and a.max_allocated == 64
    # survivors kept their contents and their identity; the param view followed its block
    assert float(ts[1].sum()) == 32.0 and float(ts[3].sum()) == 64.0 and float(p.sum()) == 48.0
    big.zero_()
    assert float(ts[1].sum()) == 32.0 and float(p.sum()) == 48.0
    ts[3].fill_(7.0)
    assert float(p.sum()) == 84.0  # still aliases the block
    a.release_tensor(ts[3])
    assert p.numel() == 0 and a.total_free == 16


def test_weight_quantization_megatron_state_dict():
    import torch
    from deepspeed_b200.runtime.weight_quantizer import WeightQuantization
    torch.manual_seed(0)
    sd = {}
    for l in range(2):
        sd[f"l{l}.attention.query_key_value.weight"] = torch.randn(96, 32)
        sd[f"l{l}.attention.dense.weight"] = torch.randn(32, 32)
        sd[f"l{l}.mlp.dense_h_to_4h.weight"] = torch.randn(128, 32)
        sd[f"l{l}.mlp.dense_4h_to_h.weight"] = torch.randn(32, 128)
        sd[f"l{l}.ln.weight"] = torch.randn(32)
    ref = {k: v.clone() for k, v in sd.items()}
    wq = WeightQuantization(mlp_extra_grouping=True)
    out, scales = wq.sd_quantize_megatron(sd, 8, 4)
    assert out["l0.ln.weight"].dtype == torch.float32 and out["l0.attention.dense.weight"].dtype == torch.int8
    assert scales.shape == (2, 4, 8)  # layers x [qkv, dense, h4h, 4hh] x widest row (MLP rows use 2x groups)
    inv = scales[0, 1, :4]  # attention.dense of layer 0: 4 groups
    deq = (out["l0.attention.dense.weight"].float().reshape(4, -1) * inv[:, None]).reshape(32, 32)
    assert (deq - ref["l0.attention.dense.weight"]).abs().max() < 0.05
    halves = wq.merge_scales_split(2)
    assert len(halves) == 2 and halves[0].shape == (2, 4, 4)
    assert wq.is_qkv(ref["l0.attention.query_key_value.weight"]) and wq.is_mlp(ref["l0.mlp.dense_h_to_4h.weight"])


# ---- engine features end to end: dataloader from `training_data`, no_sync, ZeRO-3 16-bit export, progressive layer drop,
# ---- sparse embedding gradients, wall-clock breakdown + TensorBoard / CSV monitors ------------------------------------------
import os as _os  # noqa: E402

from torch import nn  # noqa: E402
from tests.unit.simple_model import SimpleModel, base_config, make_batch  # noqa: E402,F811

os = _os


def _engine_feature_worker(which, d):
    import deepspeed_b200 as ds
    r, w = ds.comm.get_rank(), ds.comm.get_world_size()
    torch.manual_seed(0)
    if which == "io":
        xs = torch.randn(32, 8); ys = torch.randint(0, 4, (32,))
        class DS(torch.utils.data.Dataset):
            def __len__(self): return 32
            def __getitem__(self, i): return xs[i], ys[i]
        cfg = base_config(1, "fp32", 1, 0.0)
        eng, opt, loader, _ = ds.initialize(model=SimpleModel(), config=cfg, training_data=DS())
        assert loader is not None
        n = 0
        for x, y in loader:
            eng.backward(eng(x, y)); eng.step(); n += 1
        print("io batches", n, len(loader))
        assert n == len(loader) and n == 32 // (4 * w)
    elif which == "no_sync":
        cfg = base_config(1, "fp32", 2, 0.0)  # gas 2
        eng, *_ = ds.initialize(model=SimpleModel(), config=cfg)
        g = torch.Generator().manual_seed(1)
        x, y = make_batch(w, 4, g)
        with eng.no_sync():
            eng.backward(eng(x[r*4:(r+1)*4], y[r*4:(r+1)*4]))
        eng.step()
        eng.backward(eng(x[r*4:(r+1)*4], y[r*4:(r+1)*4])); eng.step()
        assert eng.global_steps == 1
    elif which == "save16":
        cfg = base_config(3, "bf16", 1, 0.0)
        cfg["zero_optimization"]["stage3_gather_16bit_weights_on_model_save"] = True
        eng, *_ = ds.initialize(model=SimpleModel(), config=cfg)
        g = torch.Generator().manual_seed(1)
        x, y = make_batch(w, 4, g)
        eng.backward(eng(x[r*4:(r+1)*4].bfloat16(), y[r*4:(r+1)*4])); eng.step()
        ok = eng.save_16bit_model(d, "model.bin")
        assert ok
        if r == 0:
            sd = torch.load(os.path.join(d, "model.bin"))
            ref = SimpleModel()
            missing = ref.load_state_dict(sd, strict=True)
            print("save16 keys", list(sd)[:3], next(iter(sd.values())).dtype)
            assert next(iter(sd.values())).dtype == torch.bfloat16
    elif which == "pld":
        cfg = base_config(0, "fp32", 1, 0.0)
        cfg["progressive_layer_drop"] = {"enabled": True, "theta": 0.5, "gamma": 0.01}
        class M(nn.Module):
            def __init__(self):
                super().__init__(); self.l = nn.Linear(8, 4); self.seen = []
            def forward(self, x, y, progressive_layer_drop=False, pld_theta=None):
                self.seen.append((progressive_layer_drop, pld_theta))
                return nn.functional.cross_entropy(self.l(x), y)
        m = M()
        eng, *_ = ds.initialize(model=m, config=cfg)
        g = torch.Generator().manual_seed(1)
        for _ in range(3):
            x, y = make_batch(1, 4, g)
            eng.backward(eng(x, y)); eng.step()
        print("pld seen", m.seen)
        assert m.seen[0][0] is True and m.seen[-1][1] < m.seen[0][1] <= 1.0
        assert abs(eng.get_pld_theta() - m.seen[-1][1]) < 0.1
    elif which == "sparse":
        cfg = base_config(0, "fp32", 1, 0.0)
        cfg["sparse_gradients"] = True
        cfg["optimizer"] = {"type": "SGD", "params": {"lr": 0.1}}
        class M(nn.Module):
            def __init__(self):
                super().__init__(); self.e = nn.Embedding(50, 8, sparse=True); self.l = nn.Linear(8, 4)
            def forward(self, ids, y): return nn.functional.cross_entropy(self.l(self.e(ids).mean(1)), y)
        torch.manual_seed(0)
        m = M(); import copy; ref = copy.deepcopy(m)
        eng, *_ = ds.initialize(model=m, config=cfg)
        ropt = torch.optim.SGD(ref.parameters(), lr=0.1)
        g = torch.Generator().manual_seed(1)
        for _ in range(3):
            ids = torch.randint(0, 50, (4 * w, 5), generator=g); y = torch.randint(0, 4, (4 * w,), generator=g)
            eng.backward(eng(ids[r*4:(r+1)*4], y[r*4:(r+1)*4])); eng.step()
            ref(ids, y).backward(); ropt.step(); ropt.zero_grad()
        from deepspeed_b200.utils import safe_get_full_fp32_param
        for (n, p), (_, q) in zip(m.named_parameters(), ref.named_parameters()):
            torch.testing.assert_close(safe_get_full_fp32_param(p).cpu(), q.detach().to_dense() if q.is_sparse else q.detach(), atol=1e-5, rtol=1e-4, msg=n)
    elif which == "wcb":
        cfg = base_config(1, "fp32", 1, 0.0)
        cfg["wall_clock_breakdown"] = True; cfg["steps_per_print"] = 1
        cfg["tensorboard"] = {"enabled": True, "output_path": d, "job_name": "tb"}
        cfg["csv_monitor"] = {"enabled": True, "output_path": d, "job_name": "csv"}
        eng, *_ = ds.initialize(model=SimpleModel(), config=cfg)
        g = torch.Generator().manual_seed(1)
        for _ in range(3):
            x, y = make_batch(w, 4, g)
            eng.backward(eng(x[r*4:(r+1)*4], y[r*4:(r+1)*4])); eng.step()
        if r == 0:
            print("monitor files", [os.path.join(dp, f) for dp, _, fs in os.walk(d) for f in fs][:6])



@pytest.mark.parametrize("which", ["io", "no_sync", "save16", "pld", "sparse", "wcb"])
def test_engine_features_end_to_end(which, tmp_path):
    from tests.common import run_distributed
    run_distributed(_engine_feature_worker, 1 if which == "pld" else 2, (which, str(tmp_path)))
