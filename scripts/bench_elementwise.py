"""Achieved HBM bandwidth of the memory-bound training kernels at Llama-3-8B shapes (tokens=8192), CUDA events, L2 flushed."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from deepspeed_b200.ops.kernels import flat_ops, transformer_ops as T

d = "cuda"
flush = torch.empty(512 << 20, dtype=torch.uint8, device=d)
peak = 6555.8


def t(fn, iters=8):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(True), torch.cuda.Event(True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


rows = []


def rec(name, ms, gbytes):
    r = {"kernel": name, "ms": round(ms, 4), "GB": round(gbytes, 3), "GBps": round(gbytes / ms * 1e3, 1),
         "pct_of_copy_peak": round(gbytes / ms * 1e3 / peak * 100, 1)}
    rows.append(r)
    print(r, flush=True)


Tn, H, I = 8192, 4096, 14336
x = torch.randn(Tn, H, device=d, dtype=torch.bfloat16)
r = torch.randn(Tn, H, device=d, dtype=torch.bfloat16)
w = torch.ones(H, device=d, dtype=torch.bfloat16)
b2 = Tn * H * 2 / 1e9
rec("rmsnorm_fwd(+residual)", t(lambda: T.rms_norm(x, w, 1e-5, residual=r)), 4 * b2)
xr, rr, wr = (v.clone().requires_grad_(True) for v in (x, r, w))
y, s = T.rms_norm(xr, wr, 1e-5, residual=rr)
gy, gs = torch.randn_like(y), torch.randn_like(s)


def nb():
    xr.grad = rr.grad = wr.grad = None
    torch.autograd.backward([y, s], [gy, gs], retain_graph=True)


rec("rmsnorm_bwd(+residual grad)", t(nb), 4 * b2)  # reads x_sum, dy, dres; writes dx (dres aliases dx)
gu = torch.randn(Tn, 2 * I, device=d, dtype=torch.bfloat16)
bg = Tn * I * 2 / 1e9
rec("swiglu_fwd", t(lambda: T.gated_act(gu, "silu")), 3 * bg)
gur = gu.clone().requires_grad_(True)
o = T.gated_act(gur, "silu")
go = torch.randn_like(o)


def gb():
    gur.grad = None
    o.backward(go, retain_graph=True)


rec("swiglu_bwd", t(gb), 5 * bg)
table = T.RotaryTable(128, 8192, 500000.0, d)
qkv = torch.randn(Tn, 6144, device=d, dtype=torch.bfloat16)
rec("rope_qk_inplace", t(lambda: T.rope_qk_inplace(qkv, 32, 8, 128, table, None, 4096)), 2 * Tn * 40 * 128 * 2 / 1e9)
lg = torch.randn(2048, 128256, device=d, dtype=torch.bfloat16)
lab = torch.randint(0, 128256, (2048, ), device=d)
rec("softmax_xent_fwd_bwd(in place)", t(lambda: T.softmax_xent_fwd_bwd(lg, lab, 1.0)), 2 * lg.numel() * 2 / 1e9)
n = 218_112_000
p = torch.randn(n, device=d)
g = torch.randn(n, device=d, dtype=torch.bfloat16)
m, v = torch.zeros(n, device=d), torch.zeros(n, device=d)
out = torch.empty(n, device=d, dtype=torch.bfloat16)
rec("adam_flat(unit)", t(lambda: flat_ops.adam_flat(p, g, m, v, out, lr=1e-4, beta1=0.9, beta2=0.95, eps=1e-8, weight_decay=0.1,
                                                 step=3)), n * 28 / 1e9)
a, c = torch.randn(Tn * H, device=d, dtype=torch.bfloat16), torch.empty(Tn * H, device=d, dtype=torch.bfloat16)
rec("torch copy (reference)", t(lambda: c.copy_(a)), 2 * b2)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/elementwise_bw.json", "w"), indent=1)
