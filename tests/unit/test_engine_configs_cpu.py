"""Engine configurations end to end (2 gloo ranks): fp32 gradient accumulation for bf16, gradient pre-division / pre-scaling,
fp16 + ZeRO-3 overflow skipping, client optimizer + torch LR scheduler factories, frozen and unused parameters with a
checkpoint round trip.  Each of these found (or guards against) a real defect."""
import copy
import os
import tempfile

import pytest
import torch
from torch import nn

from tests.common import run_distributed
from tests.unit.simple_model import SimpleModel, base_config, make_batch


def train(eng, steps, w, r, dtype=None, seed=1):
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(steps):
        x, y = make_batch(w, 4, g)
        xx = x[r*4:(r+1)*4]
        if dtype is not None: xx = xx.to(dtype)
        l = eng(xx, y[r*4:(r+1)*4]); eng.backward(l); eng.step(); out.append(l.item())
    return out


def _config_worker(which):
    import deepspeed_b200 as ds
    from deepspeed_b200.utils import safe_get_full_fp32_param
    r, w = ds.comm.get_rank(), ds.comm.get_world_size()
    torch.manual_seed(0)
    if which == "gradacc_fp32":
        cfg = base_config(2, "bf16", 2, 1.0); cfg["data_types"] = {"grad_accum_dtype": "fp32"}
        eng, *_ = ds.initialize(model=SimpleModel(), config=cfg)
        ls = train(eng, 6, w, r, torch.bfloat16); assert ls[-1] < ls[0], ls
    elif which == "predivide":
        for extra in ({"gradient_predivide_factor": 2.0}, {"prescale_gradients": True}):
            torch.manual_seed(0)
            ref = SimpleModel(); m = copy.deepcopy(ref)
            cfg = base_config(0, "fp32", 1, 0.0); cfg.update(extra); cfg["optimizer"] = {"type": "SGD", "params": {"lr": 0.1}}
            eng, *_ = ds.initialize(model=m, config=cfg)
            ropt = torch.optim.SGD(ref.parameters(), lr=0.1)
            g = torch.Generator().manual_seed(1)
            for _ in range(3):
                x, y = make_batch(w, 4, g)
                l = eng(x[r*4:(r+1)*4], y[r*4:(r+1)*4]); eng.backward(l); eng.step()
                ref(x, y).backward(); ropt.step(); ropt.zero_grad()
            for (n, p), (_, q) in zip(m.named_parameters(), ref.named_parameters()):
                torch.testing.assert_close(safe_get_full_fp32_param(p).cpu(), q.detach(), atol=1e-5, rtol=1e-4, msg=f"{extra} {n}")
    elif which == "fp16_z3_overflow":
        cfg = base_config(3, "fp16", 1, 1.0); cfg["fp16"] = {"enabled": True, "initial_scale_power": 4, "loss_scale_window": 2, "hysteresis": 1}
        eng, *_ = ds.initialize(model=SimpleModel(), config=cfg)
        g = torch.Generator().manual_seed(1)
        x, y = make_batch(w, 4, g)
        before = [safe_get_full_fp32_param(p).clone() for p in eng.module.parameters()]
        s0 = eng.optimizer.loss_scale
        l = eng((x[r*4:(r+1)*4] * 1e6).half(), y[r*4:(r+1)*4]); eng.backward(l); eng.step()   # overflow -> skipped
        after = [safe_get_full_fp32_param(p) for p in eng.module.parameters()]
        assert all(torch.equal(a, b) for a, b in zip(before, after)), "overflow step must not change parameters"
        assert eng.optimizer.loss_scale < s0 and eng.skipped_steps == 1, (eng.optimizer.loss_scale, s0, eng.skipped_steps)
        ls = train(eng, 6, w, r, torch.float16); assert ls[-1] < ls[0] and all(l == l for l in ls), ls
    elif which == "client_opt":
        for stage in (1, 2):
            torch.manual_seed(0)
            m = SimpleModel(); ref = copy.deepcopy(m)
            opt = torch.optim.Adam(m.parameters(), lr=1e-2)
            cfg = base_config(stage, "fp32", 1, 0.0); cfg.pop("optimizer", None); cfg["zero_allow_untested_optimizer"] = True
            sched = lambda o: torch.optim.lr_scheduler.StepLR(o, step_size=2, gamma=0.5)
            eng, eopt, _, esched = ds.initialize(model=m, optimizer=opt, lr_scheduler=sched, config=cfg)
            ropt = torch.optim.Adam(ref.parameters(), lr=1e-2); rs = torch.optim.lr_scheduler.StepLR(ropt, 2, 0.5)
            g = torch.Generator().manual_seed(1)
            for _ in range(4):
                x, y = make_batch(w, 4, g)
                l = eng(x[r*4:(r+1)*4], y[r*4:(r+1)*4]); eng.backward(l); eng.step()
                ref(x, y).backward(); ropt.step(); ropt.zero_grad(); rs.step()
            assert abs(eng.get_lr()[0] - ropt.param_groups[0]["lr"]) < 1e-12, (eng.get_lr(), ropt.param_groups[0]["lr"])
            for (n, p), (_, q) in zip(m.named_parameters(), ref.named_parameters()):
                torch.testing.assert_close(safe_get_full_fp32_param(p).cpu(), q.detach(), atol=1e-5, rtol=1e-4, msg=f"stage{stage} {n}")
    elif which == "frozen_unused":
        class M(nn.Module):
            def __init__(self):
                super().__init__(); self.a = nn.Linear(8, 8); self.frozen = nn.Linear(8, 8); self.unused = nn.Linear(8, 8); self.out = nn.Linear(8, 4)
                for p in self.frozen.parameters(): p.requires_grad_(False)
            def forward(self, x, y): return nn.functional.cross_entropy(self.out(self.frozen(torch.tanh(self.a(x)))), y)
        for stage in (1, 2, 3):
            torch.manual_seed(0)
            m = M(); cfg = base_config(stage, "fp32", 1, 0.0)
            eng, *_ = ds.initialize(model=m, config=cfg)
            fr0 = safe_get_full_fp32_param(m.frozen.weight).clone(); un0 = safe_get_full_fp32_param(m.unused.weight).clone()
            gg = torch.Generator().manual_seed(1); xb, yb = make_batch(w, 4, gg); ls = []
            for _ in range(6):
                l = eng(xb[r*4:(r+1)*4], yb[r*4:(r+1)*4]); eng.backward(l); eng.step(); ls.append(l.item())
            assert ls[-1] < ls[0], (stage, ls)
            assert torch.equal(safe_get_full_fp32_param(m.frozen.weight), fr0), stage
            d = tempfile.mkdtemp() if r == 0 else None
            lst = [d]; torch.distributed.broadcast_object_list(lst, 0); d = lst[0]
            eng.save_checkpoint(d, tag="t")
            a_w = safe_get_full_fp32_param(m.a.weight).clone()
            train(eng, 1, w, r, seed=9)
            eng.load_checkpoint(d, tag="t")
            torch.testing.assert_close(safe_get_full_fp32_param(m.a.weight), a_w, msg=f"a stage {stage}")
            torch.testing.assert_close(safe_get_full_fp32_param(m.frozen.weight), fr0, msg=f"frozen stage {stage}")



@pytest.mark.parametrize("which", ["gradacc_fp32", "predivide", "fp16_z3_overflow", "client_opt", "frozen_unused"])
def test_engine_configuration(which):
    run_distributed(_config_worker, 2, (which, ), timeout=400)


def _named_sched_worker():
    import deepspeed_b200 as ds
    cfg = base_config(1, "fp32", 1, 0.0)
    cfg["scheduler"] = {"type": "StepLR", "params": {"step_size": 1, "gamma": 0.5}}
    eng, *_ = ds.initialize(model=SimpleModel(), config=cfg)
    lr0 = eng.get_lr()[0]
    g = torch.Generator().manual_seed(1)
    x, y = make_batch(1, 4, g)
    eng.backward(eng(x, y))
    eng.step()
    assert abs(eng.get_lr()[0] - 0.5 * lr0) < 1e-12


def test_named_torch_scheduler_from_config():
    """``scheduler: {type: <a torch.optim.lr_scheduler class>}`` is built against the engine's parameter groups."""
    run_distributed(_named_sched_worker, 1)


# ---- ZeRO-3 usage patterns against plain PyTorch: non-reentrant activation checkpointing, evaluation between training steps,
# ---- no prefetch / one live unit, and two forward passes feeding one loss (units are released between their backward invocations)
class _PatternNet(nn.Module):
    def __init__(self, ckpt=False):
        super().__init__()
        self.layers = nn.ModuleList([nn.Linear(8, 8) for _ in range(3)])
        self.head = nn.Linear(8, 4)
        self.ckpt = ckpt
    def body(self, x):
        from torch.utils.checkpoint import checkpoint
        for l in self.layers:
            x = checkpoint(lambda t, l=l: torch.tanh(l(t)), x, use_reentrant=False) if self.ckpt else torch.tanh(l(x))
        return x
    def forward(self, x, y=None):
        h = self.head(self.body(x))
        return h if y is None else nn.functional.cross_entropy(h, y)


def _z3_pattern_worker(which):
    import deepspeed_b200 as ds
    from deepspeed_b200.utils import safe_get_full_fp32_param
    r, w = ds.comm.get_rank(), ds.comm.get_world_size()
    torch.manual_seed(0)
    extra = {}
    if which == "no_prefetch":
        extra = {"stage3_prefetch_bucket_size": 0, "stage3_max_live_parameters": 1, "stage3_max_reuse_distance": 0}
    ref = _PatternNet(ckpt=(which == "ckpt")); m = copy.deepcopy(ref)
    cfg = base_config(3, "fp32", 1, 0.0); cfg["optimizer"] = {"type": "SGD", "params": {"lr": 0.1}}
    cfg["zero_optimization"].update(extra); cfg["zero_optimization"]["stage3_param_persistence_threshold"] = 0
    eng, *_ = ds.initialize(model=m, config=cfg)
    ropt = torch.optim.SGD(ref.parameters(), lr=0.1)
    g = torch.Generator().manual_seed(1)
    for it in range(3):
        x, y = make_batch(w, 4, g)
        xs, ys = x[r*4:(r+1)*4], y[r*4:(r+1)*4]
        if which == "two_forwards":
            # two forward passes feed one loss (siamese / contrastive style)
            a = eng(xs); b = eng(xs * 0.5)
            loss = nn.functional.cross_entropy(a + b, ys)
            ra = ref(x); rb = ref(x * 0.5); rl = nn.functional.cross_entropy(ra + rb, y)
        elif which == "eval_between":
            eng.eval()
            with torch.no_grad():
                pred = eng(xs)
            assert pred.shape == (4, 4)
            eng.train()
            loss = eng(xs, ys); rl = ref(x, y)
        else:
            loss = eng(xs, ys); rl = ref(x, y)
        eng.backward(loss); eng.step()
        rl.backward(); ropt.step(); ropt.zero_grad()
    for (n, p), (_, q) in zip(m.named_parameters(), ref.named_parameters()):
        torch.testing.assert_close(safe_get_full_fp32_param(p).cpu(), q.detach(), atol=1e-5, rtol=1e-4, msg=f"{which} {n}")



@pytest.mark.parametrize("which", ["plain", "ckpt", "two_forwards", "eval_between", "no_prefetch"])
def test_zero3_usage_patterns(which):
    run_distributed(_z3_pattern_worker, 2, (which, ))


# ---- initialize() variants: config from args / config_params, explicit parameter groups with per-group lr / weight decay ------
import argparse  # noqa: E402
import json  # noqa: E402


def _init_variants_worker(which):
    import deepspeed_b200 as ds
    from deepspeed_b200.utils import safe_get_full_fp32_param
    r, w = ds.comm.get_rank(), ds.comm.get_world_size()
    torch.manual_seed(0)
    g = torch.Generator().manual_seed(1); x, y = make_batch(w, 4, g); xs, ys = x[r*4:(r+1)*4], y[r*4:(r+1)*4]
    if which == "args_path":
        d = tempfile.mkdtemp(); p = os.path.join(d, "c.json"); json.dump(base_config(1, "fp32", 1, 0.0), open(p, "w"))
        ap = argparse.ArgumentParser(); ap = ds.add_config_arguments(ap)
        args = ap.parse_args(["--deepspeed", "--deepspeed_config", p])
        eng, *_ = ds.initialize(args=args, model=SimpleModel())
        eng.backward(eng(xs, ys)); eng.step()
        eng2, *_ = ds.initialize(model=SimpleModel(), config_params=base_config(2, "fp32", 1, 0.0))
        eng2.backward(eng2(xs, ys)); eng2.step()
    elif which == "param_groups":
        for stage in (0, 1, 2, 3):
            torch.manual_seed(0)
            m = SimpleModel(); ref = copy.deepcopy(m)
            decay = [p for n, p in m.named_parameters() if p.dim() > 1]; nodecay = [p for n, p in m.named_parameters() if p.dim() <= 1]
            rdecay = [p for n, p in ref.named_parameters() if p.dim() > 1]; rnodecay = [p for n, p in ref.named_parameters() if p.dim() <= 1]
            cfg = base_config(stage, "fp32", 1, 0.0); cfg["optimizer"] = {"type": "AdamW", "params": {"lr": 1e-2, "weight_decay": 0.1}}
            eng, *_ = ds.initialize(model=m, config=cfg, model_parameters=[{"params": decay}, {"params": nodecay, "weight_decay": 0.0, "lr": 5e-2}])
            ropt = torch.optim.AdamW([{"params": rdecay}, {"params": rnodecay, "weight_decay": 0.0, "lr": 5e-2}], lr=1e-2, weight_decay=0.1)
            gg = torch.Generator().manual_seed(1)
            for _ in range(3):
                xx, yy = make_batch(w, 4, gg)
                eng.backward(eng(xx[r*4:(r+1)*4], yy[r*4:(r+1)*4])); eng.step()
                ref(xx, yy).backward(); ropt.step(); ropt.zero_grad()
            for (n, p), (_, q) in zip(m.named_parameters(), ref.named_parameters()):
                torch.testing.assert_close(safe_get_full_fp32_param(p).cpu(), q.detach(), atol=2e-5, rtol=1e-4, msg=f"stage {stage} {n}")
            assert [round(l, 6) for l in eng.get_lr()] == [1e-2, 5e-2], eng.get_lr()



@pytest.mark.parametrize("which", ["args_path", "param_groups"])
def test_initialize_variants(which):
    run_distributed(_init_variants_worker, 2, (which, ), timeout=400)


def _tp_consolidate_worker(d):
    import copy
    import deepspeed_b200 as ds
    from deepspeed_b200 import comm as dist
    from deepspeed_b200.module_inject.layers import LinearAllreduce, LinearLayer, set_autotp_mode
    from deepspeed_b200.utils import groups
    set_autotp_mode(training=True)
    torch.manual_seed(0)
    a, b = torch.nn.Linear(16, 32), torch.nn.Linear(32, 16)
    cfg = {"train_micro_batch_size_per_gpu": 1, "optimizer": {"type": "SGD", "params": {"lr": 0.0}},
           "tensor_parallel": {"autotp_size": 2}, "zero_optimization": {"stage": 1}}
    boot, *_ = ds.initialize(model=torch.nn.Linear(2, 2), config=cfg)  # creates the TP groups
    g = groups.get_tensor_model_parallel_group()

    class TP(torch.nn.Module):

        def __init__(self):
            super().__init__()
            self.a, self.b = LinearLayer(copy.deepcopy(a), g), LinearAllreduce(copy.deepcopy(b), g)

        def forward(self, x):
            return self.b(torch.relu(self.a(x))).sum()

    model = TP()
    assert model.a.weight.shape == (16, 16) and model.b.weight.shape == (16, 16)
    # parameter-level API of the reference: gather in place, then re-partition
    w = model.a.weight
    w.gather_params([w, model.a.bias])
    assert torch.equal(w, a.weight) and torch.equal(model.a.bias, a.bias)
    w._tp_partition([w, model.a.bias])
    assert w.shape == (16, 16) and model.a.bias.shape == (16, )
    eng, *_ = ds.initialize(model=model, config=cfg)
    eng.backward(eng(torch.randn(1, 16)))
    eng.step()
    sd = eng._consolidated_16bit_state_dict()
    assert eng.save_16bit_model(d, "tp.bin")
    if dist.get_rank() == 0:
        ref = {"a.weight": a.weight, "a.bias": a.bias, "b.weight": b.weight, "b.bias": b.bias}
        on_disk = torch.load(os.path.join(d, "tp.bin"), weights_only=False)
        for k, v in ref.items():
            torch.testing.assert_close(sd[k].float(), v.detach())
            torch.testing.assert_close(on_disk[k].float(), v.detach())
    else:
        assert sd is None
    assert model.a.weight.shape == (16, 16)  # shards restored


def test_tensor_parallel_training_consolidated_state_dict(tmp_path):
    run_distributed(_tp_consolidate_worker, 2, (str(tmp_path), ))


def _basic_optimizer_worker():
    """Without ZeRO / mixed precision the reference hands the *basic* optimizer back from ``initialize`` (reference
    tests/unit/runtime/test_ds_initialize.py TestClientOptimizer / TestConfigOptimizer) and refuses (model dtype,
    gradient-accumulation dtype) pairs that have no wrapper (TestOptimizerImplementation)."""
    import deepspeed_b200 as ds
    from deepspeed_b200.ops.adam import FusedAdam
    g = torch.Generator().manual_seed(0)

    def step(eng):
        x, y = make_batch(1, 4, g)
        l = eng(x[:4], y[:4]); eng.backward(l); eng.step()
        return l.item()

    # config-named Adam -> FusedAdam; still the engine's flat optimizer (it trains)
    m = SimpleModel()
    eng, opt, _, _ = ds.initialize(model=m, model_parameters=list(m.parameters()),
                                   config={"train_batch_size": 4, "optimizer": {"type": "Adam", "params": {"lr": 1e-2}}})
    assert isinstance(opt, FusedAdam) and isinstance(opt, torch.optim.Optimizer) and opt is eng.optimizer
    l0 = step(eng); [step(eng) for _ in range(4)]
    assert step(eng) < l0 * 1.5
    # client optimizer object -> compares equal to it, is an instance of its class
    m = SimpleModel()
    client = torch.optim.Adam(m.parameters(), lr=1e-2)
    eng, opt, _, _ = ds.initialize(model=m, optimizer=client, config={"train_batch_size": 4})
    assert opt == client and isinstance(opt, torch.optim.Adam) and not (opt == torch.optim.Adam(m.parameters()))
    step(eng)
    # optimizer factory -> instance of what the factory builds
    m = SimpleModel()
    eng, opt, _, _ = ds.initialize(model=m, model_parameters=list(m.parameters()), config={"train_batch_size": 4},
                                   optimizer=lambda params: torch.optim.AdamW(params, lr=1e-2))
    assert isinstance(opt, torch.optim.AdamW)
    step(eng)
    # unsupported dtype pairs without ZeRO are refused; the same pairs are fine with ZeRO
    for bf16, gad, ok in ((True, "bf16", False), (False, "bf16", False), (True, "fp32", True), (False, "fp32", True),
                          (True, None, True)):
        cfg = {"train_batch_size": 4, "bf16": {"enabled": bf16}, "data_types": {"grad_accum_dtype": gad},
               "optimizer": {"type": "Adam", "params": {"lr": 1e-3}}}
        m = SimpleModel()
        if ok:
            ds.initialize(model=m, model_parameters=list(m.parameters()), config=cfg)
        else:
            with pytest.raises(NotImplementedError):
                ds.initialize(model=m, model_parameters=list(m.parameters()), config=cfg)
            cfg["zero_optimization"] = {"stage": 1}
            m = SimpleModel()
            ds.initialize(model=m, model_parameters=list(m.parameters()), config=cfg)


def test_basic_optimizer_class_and_dtype_pairs():
    run_distributed(_basic_optimizer_worker, 1, timeout=300)
