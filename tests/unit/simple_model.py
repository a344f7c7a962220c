"""Small fixtures shared by the CPU-tier tests (role: reference tests/unit/simple_model.py)."""
import torch
from torch import nn


class Block(nn.Module):

    def __init__(self, d):
        super().__init__()
        self.a = nn.Linear(d, d)
        self.n = nn.LayerNorm(d)

    def forward(self, x):
        return x + torch.relu(self.a(self.n(x)))


class SimpleModel(nn.Module):

    def __init__(self, d=32, nlayers=3, nclass=4, in_dim=8):
        super().__init__()
        self.emb = nn.Linear(in_dim, d)
        self.layers = nn.ModuleList([Block(d) for _ in range(nlayers)])
        self.head = nn.Linear(d, nclass)

    def forward(self, x, y):
        h = self.emb(x)
        for layer in self.layers:
            h = layer(h)
        return nn.functional.cross_entropy(self.head(h).float(), y)


def make_batch(world, per_rank, gen, in_dim=8, nclass=4):
    x = torch.randn(per_rank * world, in_dim, generator=gen)
    y = torch.randint(0, nclass, (per_rank * world, ), generator=gen)
    return x, y


def base_config(stage, dtype="fp32", gas=1, clip=0.0, lr=1e-2, opt="AdamW", extra=None):
    cfg = {
        "train_micro_batch_size_per_gpu": 4,
        "gradient_accumulation_steps": gas,
        "optimizer": {"type": opt, "params": {"lr": lr}},
        "zero_optimization": {"stage": stage},
        "gradient_clipping": clip,
    }
    if stage == 3:  # tiny test models would otherwise be 'persistent' and never exercise gather/release
        cfg["zero_optimization"]["stage3_param_persistence_threshold"] = 0
    if opt.lower() in ("adamw", "adam"):
        cfg["optimizer"]["params"]["weight_decay"] = 0.01
    if dtype == "bf16":
        cfg["bf16"] = {"enabled": True}
    if dtype == "fp16":
        cfg["fp16"] = {"enabled": True, "initial_scale_power": 8}
    if extra:
        for k, v in extra.items():
            if isinstance(v, dict) and isinstance(cfg.get(k), dict):
                cfg[k].update(v)
            else:
                cfg[k] = v
    return cfg
