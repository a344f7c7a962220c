"""Native async I/O engine, swap buffers, NVMe-backed optimizer state, ZeRO + NVMe offload vs no offload."""
import os

import pytest
import torch

from tests.common import run_distributed
from tests.unit.simple_model import SimpleModel, base_config, make_batch


def test_aio_roundtrip(tmp_path):
    from deepspeed_b200.ops.aio import aio_handle, file_size
    h = aio_handle(block_size=1 << 16, queue_depth=4, intra_op_parallelism=2)
    assert h.get_block_size() == 1 << 16 and h.get_queue_depth() == 4 and h.get_intra_op_parallelism() == 2
    src = h.new_cpu_locked_tensor(300_001, torch.empty(0, dtype=torch.float32))
    src.copy_(torch.randn(300_001))
    path = str(tmp_path / "a.swp")
    h.sync_pwrite(src, path)
    assert file_size(path) == src.numel() * 4
    dst = torch.empty_like(src)
    h.async_pread(dst, path)
    assert h.wait() == 1
    assert torch.equal(src, dst)
    # offset I/O (unaligned -> buffered fallback inside the engine)
    part = torch.empty(1000, dtype=torch.float32)
    h.sync_pread(part, path, file_offset=4 * 777)
    assert torch.equal(part, src[777:1777])
    h.free_cpu_locked_tensor(src)


def test_swap_buffer_pool_and_async_swapper(tmp_path):
    from deepspeed_b200.ops.aio import aio_handle
    from deepspeed_b200.runtime.swap_tensor import AsyncTensorSwapper, SwapBufferManager, SwapBufferPool
    h = aio_handle()
    mgr = SwapBufferManager(num_elems=4096, count=3, dtype=torch.float32)
    bufs = mgr.allocate(4096, 2, torch.float32)
    pool = SwapBufferPool(bufs)
    ts = [torch.randn(1000), torch.randn(3000), torch.randn(2000)]
    paths = [str(tmp_path / f"t{i}.swp") for i in range(3)]
    for t, p in zip(ts, paths):
        st, ct = pool.insert_tensor(t, p, 1024 * ((t.numel() + 1023) // 1024))
        assert st is not None
    pool.swap_out(h)
    pool.reset()
    for t, p in zip(ts, paths):
        pool.allocate_tensor(t.numel(), p, 1024 * ((t.numel() + 1023) // 1024))
    pool.swap_in(h)
    for t, c in zip(ts, pool.get_compute_tensors()):
        assert torch.equal(t, c)
    mgr.free(bufs)
    sw = AsyncTensorSwapper(h, numel_alignment=256)
    sw.add_buffers(mgr.allocate(4096, 2, torch.float32))
    big = [torch.randn(3000) for _ in range(5)]
    bp = [str(tmp_path / f"g{i}.swp") for i in range(5)]
    sw.swap_out_tensors(big, bp)
    sw.release_buffers()
    for t, p in zip(big, bp):
        back = torch.empty(3072)
        h.sync_pread(back, p)
        assert torch.equal(back[:3000], t)


def test_swapped_flat_state(tmp_path):
    from deepspeed_b200.ops.aio import aio_handle
    from deepspeed_b200.runtime.swap_tensor import SwappedFlatState
    st = SwappedFlatState("exp_avg", 10_000, torch.float32, str(tmp_path), aio_handle(), window_elems=3000, n_windows=3)
    ref = torch.zeros(10_000)
    for s in range(0, 10_000, 3000):
        e = min(s + 3000, 10_000)
        if e < 10_000:
            st.prefetch(e, min(e + 3000, 10_000))
        w = st[s:e]
        w.add_(torch.arange(s, e, dtype=torch.float32))
        ref[s:e] += torch.arange(s, e, dtype=torch.float32)
    st.flush()
    assert torch.equal(st.detach(), ref)
    st.copy_(ref * 2)
    assert torch.equal(st[100:200], ref[100:200] * 2)


def _train(mode, nvme_dir, out):
    import deepspeed_b200 as ds
    torch.manual_seed(0)
    cfg = base_config(2, "bf16", 1, 1.0)
    if mode == "cpu":
        cfg["zero_optimization"]["offload_optimizer"] = {"device": "cpu"}
    elif mode == "nvme":
        cfg["zero_optimization"]["offload_optimizer"] = {"device": "nvme", "nvme_path": nvme_dir, "b200_swap_window": 200,
                                                          "pipeline_read": True, "pipeline_write": True}
        cfg["aio"] = {"block_size": 65536, "queue_depth": 4}
    eng, *_ = ds.initialize(model=SimpleModel(), config=cfg)
    g = torch.Generator().manual_seed(1)
    for _ in range(4):
        x, y = make_batch(1, 4, g)
        loss = eng(x.bfloat16(), y)
        eng.backward(loss)
        eng.step()
    from deepspeed_b200.utils import safe_get_full_fp32_param, safe_get_full_optimizer_state
    sd = {n: safe_get_full_fp32_param(p).cpu() for n, p in eng.module.named_parameters()}
    torch.save(sd, out)
    if mode == "nvme":
        from deepspeed_b200.runtime.swap_tensor.optimizer_utils import SwappedFlatState
        from deepspeed_b200.utils import safe_set_full_fp32_param
        zo = eng.optimizer
        assert zo.state_swapper is not None
        assert isinstance(zo.master, SwappedFlatState), "the fp32 master weights must be swapped too, not host resident"
        files = [f for _, _, fs in os.walk(nvme_dir) for f in fs]
        assert "fp32_master.swp" in files and any(f.startswith("exp_avg") for f in files), files
        # debug / checkpoint surface on top of the file-backed arrays (ranges larger than the 200-element window)
        p0 = next(iter(eng.module.parameters()))
        before = safe_get_full_fp32_param(p0).clone()
        safe_set_full_fp32_param(p0, before + 1.0)
        torch.testing.assert_close(safe_get_full_fp32_param(p0), before + 1.0)
        safe_set_full_fp32_param(p0, before)
        m = safe_get_full_optimizer_state(p0, "exp_avg")
        assert m is not None and m.shape == p0.shape and m.abs().sum() > 0
        ck = os.path.join(nvme_dir, "ckpt")
        eng.save_checkpoint(ck, tag="t")
        x, y = make_batch(1, 4, g)
        eng.backward(eng(x.bfloat16(), y))
        eng.step()
        moved = safe_get_full_fp32_param(p0).clone()
        assert not torch.equal(moved, before)
        eng.load_checkpoint(ck, tag="t")
        torch.testing.assert_close(safe_get_full_fp32_param(p0), before)  # master restored from the checkpoint


def test_nvme_offload_matches_cpu_offload(tmp_path):
    a, b = str(tmp_path / "cpu.pt"), str(tmp_path / "nvme.pt")
    run_distributed(_train, 1, ("cpu", "", a))
    run_distributed(_train, 1, ("nvme", str(tmp_path / "nvme"), b))
    sa, sb = torch.load(a), torch.load(b)
    for k in sa:
        torch.testing.assert_close(sa[k], sb[k], atol=1e-6, rtol=1e-6)


@pytest.mark.parametrize("kind", ["partitioned", "pipelined"])
def test_named_optimizer_swappers(tmp_path, kind):
    """The two reference swapper classes drive the same flat NVMe-backed state, with and without read-ahead."""
    from types import SimpleNamespace
    from deepspeed_b200.runtime.swap_tensor.partitioned_optimizer_swapper import PartitionedOptimizerSwapper
    from deepspeed_b200.runtime.swap_tensor.pipelined_optimizer_swapper import PipelinedOptimizerSwapper
    cls = PartitionedOptimizerSwapper if kind == "partitioned" else PipelinedOptimizerSwapper
    sw = cls(SimpleNamespace(b200_swap_window=1000, buffer_count=4), {"block_size": 65536, "queue_depth": 4}, str(tmp_path))
    assert sw.pipeline == (kind == "pipelined")
    flat = SimpleNamespace(state_names=["exp_avg", "exp_avg_sq"], state={})
    sw.wrap(flat, 3500)
    ref = torch.zeros(3500)
    for s in range(0, 3500, 1000):
        e = min(s + 1000, 3500)
        nxt = (e, min(e + 1000, 3500)) if e < 3500 else None
        win = sw.swap_in_optimizer_state(flat, s, e, nxt) if kind == "pipelined" else sw.swap_in_optimizer_state(flat, s, e)
        win["exp_avg"].add_(1.5)
        win["exp_avg_sq"].add_(torch.arange(s, e, dtype=torch.float32))
        ref[s:e] = torch.arange(s, e, dtype=torch.float32)
    sw.swap_out_optimizer_state(flat, async_swap=False)
    assert torch.equal(flat.state["exp_avg"].detach(), torch.full((3500, ), 1.5))
    assert torch.equal(flat.state["exp_avg_sq"].detach(), ref)


def test_per_parameter_optimizer_swapper(tmp_path):
    import torch
    from deepspeed_b200.runtime.swap_tensor.optimizer_utils import (FlattenedTensorSwapInfo, OptimizerStateSwapInfo,
                                                                      OptimizerSwapper)
    from deepspeed_b200.runtime.swap_tensor.pipelined_optimizer_swapper import OptimizerSwapOp
    p = torch.nn.Parameter(torch.randn(512, 1024))  # 2 MiB fp32: above the aio threshold
    small = torch.nn.Parameter(torch.randn(8))
    opt = torch.optim.Adam([p, small], lr=1e-2)
    p.grad, small.grad = torch.randn_like(p), torch.randn_like(small)
    opt.step()
    sw = OptimizerSwapper(None, {"block_size": 1 << 20, "queue_depth": 8, "intra_op_parallelism": 1, "single_submit": False,
                                 "overlap_events": True, "use_gds": False}, str(tmp_path), opt, p.numel(), "cpu",
                          torch.float32, None)
    assert sw.swappable_tensor(param=p) and not sw.swappable_tensor(param=small)
    want = [p.detach().clone(), opt.state[p]["exp_avg"].clone(), opt.state[p]["exp_avg_sq"].clone()]
    sw.swap_out_optimizer_state(p)
    assert p.numel() == 0 and opt.state[p]["exp_avg"].numel() == 0
    g = torch.randn(512 * 1024)
    half = g.numel() // 2
    sw.swap_out_gradients(p, [0, half, g.numel() - 16], [g[:half], g[half:g.numel() - 16], g[-16:]])
    info = sw._get_param_swap_info(p)
    assert isinstance(info, OptimizerStateSwapInfo) and info.has_gradients() and len(info.unswapped_gradients) == 2
    assert all(isinstance(x, FlattenedTensorSwapInfo) for x in info.swapped_gradients.values())
    sw.swap_in_optimizer_state(p)
    assert torch.equal(p.detach(), want[0]) and torch.equal(opt.state[p]["exp_avg"], want[1])
    assert torch.equal(opt.state[p]["exp_avg_sq"], want[2]) and torch.equal(p.grad.reshape(-1), g)
    opt.step()  # the optimizer keeps working on the swapped-in storage
    op = OptimizerSwapOp(sw.aio_handle, True, info, [], [], 3)
    assert op.is_parameter(p) and not op.is_parameter(small)
    op.wait()
    assert not op.wait_required


def test_nvme_benchmark_schedules(tmp_path):
    import types
    from deepspeed_b200.nvme import ds_aio_basic as B, ds_aio_handle as H, perf_run_sweep as S
    args = types.SimpleNamespace(mapping_list=[(0, str(tmp_path))], io_size=1 << 20, loops=2, block_size=1 << 18, queue_depth=4,
                                 single_submit=False, sequential_requests=False, io_parallel=2, use_gds=False, gpu=False,
                                 validate=False, multi_process=1)
    for mod in (B, H):
        for read_op in (False, True):
            sched = mod.get_schedule(args, read_op)
            assert set(sched) == {"pre", "main", "post"}
            ctxt = sched["pre"]((args, 0))
            for _ in range(args.loops):
                sched["main"]((args, 0, ctxt))
            assert ctxt["elapsed_sec"] > 0 and ctxt["num_bytes"] == 1 << 20
            sched["post"]((args, 0, ctxt))
    assert H.get_schedule(args, True)["main"] is H.main_parallel_read
    args.io_parallel = 1
    assert H.get_schedule(args, True)["main"] is H.main_handle_read
    jobs = S.create_perf_jobs("read", str(tmp_path), [["--block_size", "1M", "--queue_depth", "8"]])
    assert len(jobs) == 1 and "--read" in jobs[0].cmd() and jobs[0].output_file.endswith(".txt")
    assert S.async_io_setup() in (True, False) and S.script_path().endswith("nvme")


def test_param_swapper_param_object_api(tmp_path):
    import types
    import torch
    from deepspeed_b200.runtime.swap_tensor.partitioned_param_swapper import AsyncPartitionedParameterSwapper, PartitionedParamStatus
    oc = types.SimpleNamespace(nvme_path=str(tmp_path), buffer_size=4096, buffer_count=3)
    sw = AsyncPartitionedParameterSwapper(oc, torch.float32, aio_config={"block_size": 1 << 20, "queue_depth": 8,
                                                                        "intra_op_parallelism": 1, "single_submit": False,
                                                                        "overlap_events": True, "use_gds": False})
    params = [types.SimpleNamespace(ds_id=i) for i in range(2)]
    new = [torch.arange(1000.0), torch.arange(2000.0) * 2]
    sw.reserve_partitioned_swap_space([t.numel() for t in new])
    sw.swap_out_partitioned_params(params, new)
    assert all(sw.status(p.ds_id) == PartitionedParamStatus.NOT_AVAILABLE for p in params)
    assert sw.get_path(params[1], must_exist=True).endswith("1_param.tensor.swp")
    dst = torch.zeros(2000)
    sw.swap_into_buffer(params[1], dst)
    assert torch.equal(dst, new[1]) and sw.status(1) == PartitionedParamStatus.AVAILABLE
    bufs = sw.reserve_available_buffers()
    assert len(bufs) == 3 and sw.available_swap_in_buffers() == 0
    sw.release_reserved_buffers()
    assert sw.available_swap_in_buffers() == 3
    got = sw.swap_in([0], async_op=False)[0]
    assert torch.equal(got, new[0])
    sw.remove_partition_and_release_buffers([params[0]])
    assert sw.available_swap_in_buffers() == 3


def _param_tier(mode, nvme_dir, out):
    """ZeRO-3 with the parameter shards on the host (pinned) or in NVMe swap files: identical training trajectory."""
    import deepspeed_b200 as ds
    from deepspeed_b200.utils import safe_get_full_fp32_param
    torch.manual_seed(0)
    cfg = base_config(3, "bf16", 1, 1.0)
    z = cfg["zero_optimization"]
    z["stage3_param_persistence_threshold"] = 0
    z["offload_optimizer"] = {"device": "cpu"}
    if mode == "cpu":
        z["offload_param"] = {"device": "cpu"}
    else:
        z["offload_param"] = {"device": "nvme", "nvme_path": nvme_dir, "buffer_count": 3}
        cfg["aio"] = {"block_size": 65536, "queue_depth": 4}
    eng, *_ = ds.initialize(model=SimpleModel(), config=cfg)
    r, w = ds.comm.get_rank(), ds.comm.get_world_size()
    if mode == "nvme":
        from deepspeed_b200.runtime.swap_tensor.optimizer_utils import SwappedFlatState
        assert isinstance(eng.optimizer.lp_arena, SwappedFlatState)
    g = torch.Generator().manual_seed(1)
    for _ in range(4):
        x, y = make_batch(w, 4, g)
        loss = eng(x[r * 4:(r + 1) * 4].bfloat16(), y[r * 4:(r + 1) * 4])
        eng.backward(loss)
        eng.step()
    sd = {n: safe_get_full_fp32_param(p).cpu() for n, p in eng.module.named_parameters()}
    if mode == "nvme":  # checkpoint round trip on top of the file-backed shards
        ck = os.path.join(nvme_dir, "ckpt")
        eng.save_checkpoint(ck, tag="t")
        x, y = make_batch(w, 4, g)
        eng.backward(eng(x[r * 4:(r + 1) * 4].bfloat16(), y[r * 4:(r + 1) * 4]))
        eng.step()
        eng.load_checkpoint(ck, tag="t")
        for n, p in eng.module.named_parameters():
            torch.testing.assert_close(safe_get_full_fp32_param(p).cpu(), sd[n])
        x, y = make_batch(w, 4, torch.Generator().manual_seed(5))
        assert torch.isfinite(eng(x[r * 4:(r + 1) * 4].bfloat16(), y[r * 4:(r + 1) * 4]))
    if r == 0:
        torch.save(sd, out)
        if mode == "nvme":
            files = [f for _, _, fs in os.walk(nvme_dir) for f in fs]
            assert "lp_params.swp" in files, files


def test_nvme_param_tier_matches_host_param_tier(tmp_path):
    a, b = str(tmp_path / "cpu.pt"), str(tmp_path / "nvme.pt")
    run_distributed(_param_tier, 2, ("cpu", "", a))
    run_distributed(_param_tier, 2, ("nvme", str(tmp_path / "nvme"), b))
    sa, sb = torch.load(a), torch.load(b)
    for k in sa:
        torch.testing.assert_close(sa[k], sb[k], atol=1e-6, rtol=1e-6)
