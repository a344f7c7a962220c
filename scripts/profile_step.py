"""Kernel-time breakdown of one training step (torch.profiler / CUPTI): prints the top kernels by
total device time.  Usage: python scripts/profile_step.py [--layers N] [--model llama3-8b] ..."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import deepspeed_b200 as ds
from deepspeed_b200.models.llama import LlamaForCausalLM, llama_config

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="llama3-8b")
ap.add_argument("--layers", type=int, default=None)
ap.add_argument("--seq", type=int, default=4096)
ap.add_argument("--micro-batch", type=int, default=2)
ap.add_argument("--ckpt", type=int, default=0)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--out", default="gpurun_out/step_profile.txt")
a = ap.parse_args()
over = {} if a.layers is None else {"num_hidden_layers": a.layers}
cfg = llama_config(a.model, checkpoint_layers=a.ckpt, **over)
torch.set_default_dtype(torch.bfloat16)
with torch.device("cuda"):
    model = LlamaForCausalLM(cfg)
torch.set_default_dtype(torch.float32)
eng, _, _, _ = ds.initialize(model=model, config={
    "train_micro_batch_size_per_gpu": a.micro_batch, "bf16": {"enabled": True},
    "optimizer": {"type": "AdamW", "params": {"lr": 1e-5, "weight_decay": 0.1}},
    "zero_optimization": {"stage": 3}, "steps_per_print": 10**9})
ids = torch.randint(0, cfg.vocab_size, (a.micro_batch, a.seq), device="cuda")


def step():
    loss = eng(ids, labels=ids)
    eng.backward(loss)
    eng.step()


for _ in range(2):
    step()
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True) as prof:
    for _ in range(a.steps):
        step()
    torch.cuda.synchronize()
tab = prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=90)
os.makedirs(os.path.dirname(a.out), exist_ok=True)
with open(a.out, "w") as f:
    f.write(f"# {a.model} layers={cfg.num_hidden_layers} seq={a.seq} mb={a.micro_batch} steps={a.steps}\n")
    f.write(tab)
    # who launches the generic ATen element-wise kernels (fills / copies / adds)?  grouped by input shape
    rows = [r for r in prof.key_averages(group_by_input_shape=True)
            if any(k in r.key for k in ("fill_", "zero_", "copy_", "aten::add", "aten::mul", "aten::cat", "index"))
            and r.device_time_total > 200]
    rows.sort(key=lambda r: -r.device_time_total)
    f.write("\n\n# generic ATen element-wise ops by input shape (device time over the profiled steps)\n")
    for r in rows[:40]:
        f.write(f"{r.key:40s} calls={r.count:5d} cuda_ms={r.device_time_total / 1e3:9.3f} shapes={r.input_shapes}\n")
print(tab[-6000:])
