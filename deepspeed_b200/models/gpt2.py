"""GPT-2 (BASELINE.json config 1: GPT-2 small ZeRO-1 on CPU/gloo, the no-GPU plumbing config)."""
from dataclasses import dataclass

import torch
import torch.nn.functional as F
from torch import nn

from deepspeed_b200.ops.kernels import transformer_ops as T
from deepspeed_b200.ops.attention import causal_attention


@dataclass
class GPT2Config:
    vocab_size: int = 50257
    n_positions: int = 1024
    n_embd: int = 768
    n_layer: int = 12
    n_head: int = 12
    layer_norm_epsilon: float = 1e-5
    initializer_range: float = 0.02
    tie_word_embeddings: bool = True


GPT2_PRESETS = {
    "gpt2-small": dict(),
    "gpt2-medium": dict(n_embd=1024, n_layer=24, n_head=16),
    "gpt2-tiny": dict(vocab_size=512, n_positions=128, n_embd=64, n_layer=2, n_head=4),
}


def gpt2_config(name, **over):
    d = dict(GPT2_PRESETS[name])
    d.update(over)
    return GPT2Config(**d)


class GPT2Block(nn.Module):

    def __init__(self, cfg):
        super().__init__()
        h = cfg.n_embd
        self.ln_1 = nn.LayerNorm(h, eps=cfg.layer_norm_epsilon)
        self.c_attn = nn.Linear(h, 3 * h)
        self.c_proj = nn.Linear(h, h)
        self.ln_2 = nn.LayerNorm(h, eps=cfg.layer_norm_epsilon)
        self.c_fc = nn.Linear(h, 4 * h)
        self.c_proj2 = nn.Linear(4 * h, h)
        self.n_head = cfg.n_head

    def forward(self, x):
        B, S, H = x.shape
        h = T.layer_norm(x, self.ln_1.weight, self.ln_1.bias, self.ln_1.eps)
        qkv = self.c_attn(h)
        a = causal_attention(qkv.view(B * S, 3 * H), B, S, self.n_head, self.n_head, H // self.n_head, None, None)
        x = x + self.c_proj(a.view(B, S, H))
        h = T.layer_norm(x, self.ln_2.weight, self.ln_2.bias, self.ln_2.eps)
        return x + self.c_proj2(T.bias_act(F.linear(h, self.c_fc.weight), self.c_fc.bias, "gelu_tanh"))


class GPT2LMHeadModel(nn.Module):

    def __init__(self, cfg: GPT2Config):
        super().__init__()
        self.cfg = cfg
        self.wte = nn.Embedding(cfg.vocab_size, cfg.n_embd)
        self.wpe = nn.Embedding(cfg.n_positions, cfg.n_embd)
        self.h = nn.ModuleList([GPT2Block(cfg) for _ in range(cfg.n_layer)])
        self.ln_f = nn.LayerNorm(cfg.n_embd, eps=cfg.layer_norm_epsilon)
        self.lm_head = nn.Linear(cfg.n_embd, cfg.vocab_size, bias=False)
        if cfg.tie_word_embeddings:
            self.lm_head.weight = self.wte.weight
        self.apply(self._init)

    def _init(self, m):
        if isinstance(m, (nn.Linear, nn.Embedding)) and m.weight.numel():
            nn.init.normal_(m.weight, std=self.cfg.initializer_range)
            if isinstance(m, nn.Linear) and m.bias is not None:
                nn.init.zeros_(m.bias)

    def forward(self, input_ids, labels=None):
        B, S = input_ids.shape
        pos = torch.arange(S, device=input_ids.device)
        x = self.wte(input_ids) + self.wpe(pos)[None]
        for blk in self.h:
            x = blk(x)
        x = T.layer_norm(x, self.ln_f.weight, self.ln_f.bias, self.ln_f.eps)
        logits = self.lm_head(x)
        if labels is None:
            return logits
        return T.cross_entropy(logits[:, :-1].reshape(-1, logits.shape[-1]), labels[:, 1:].reshape(-1))
