"""FPDT / Ulysses-Offload: chunked sequence pipeline with online-softmax merging and optional host offload.

Parity target: reference ``sequence/fpdt_layer.py`` (``update_out_and_lse :58``, ``FPDT_InputConstruct :79``,
``_FPDTGPUAttentionImpl_ :134``, ``_FPDTGPUOffloadingAttentionImpl_ :510``, ``SequenceChunk :462``,
``FPDT_Attention :971``, ``FPDT_FFN :1056``, ``FPDT_LogitsLoss :1137``).

The local sequence is split into ``num_chunks`` chunks.  For query chunk ``i`` the projections are computed,
Ulysses all-to-all exchanges heads for sequence, and the chunk attends to key/value chunks ``0..i`` (causal
only on the diagonal pair); partial outputs are merged with the running log-sum-exp.  With ``offloading=True``
processed K/V chunks live in pinned host memory and are prefetched back on a side stream one pair ahead
(double buffering), so device memory holds O(chunk) activations for arbitrarily long sequences.  The
backward is obtained by autograd over the per-pair attention calls wrapped in activation checkpoints, i.e.
each (q-chunk, kv-chunk) pair is recomputed exactly like the reference's manual backward.
"""
import math
from typing import Optional

import torch
import torch.nn.functional as F

from deepspeed_b200 import comm as dist
from deepspeed_b200.sequence.layer import single_all_to_all, _SeqAllToAll


def update_out_and_lse(out, lse, block_out, block_lse):
    """Merge two partial attention results over disjoint key sets (all fp32, lse shape [..., S, 1])."""
    if out is None:
        return block_out.float(), block_lse.float()
    new_lse = lse + F.softplus(block_lse - lse)  # log(exp(lse) + exp(block_lse)), stable
    out = torch.exp(lse - new_lse) * out + torch.exp(block_lse - new_lse) * block_out.float()
    return out, new_lse


def _attn_with_lse(q, k, v, causal, scale):
    """q,k,v [B,H,S,D] -> (out fp32, lse fp32 [B,H,S,1]).  GQA by head repetition."""
    if k.shape[1] != q.shape[1]:
        rep = q.shape[1] // k.shape[1]
        k, v = k.repeat_interleave(rep, dim=1), v.repeat_interleave(rep, dim=1)
    s = torch.matmul(q.float(), k.float().transpose(-1, -2)) * scale
    if causal:
        Sq, Sk = s.shape[-2:]
        s = s.masked_fill(~torch.ones(Sq, Sk, dtype=torch.bool, device=s.device).tril_(diagonal=Sk - Sq), float("-inf"))
    lse = torch.logsumexp(s, dim=-1, keepdim=True)
    return torch.matmul(torch.exp(s - lse), v.float()), lse


class SequenceChunk:
    """A K/V (or Q) chunk that may live on the host; ``load_to_gpu`` is asynchronous on a side stream."""

    def __init__(self, chunk: torch.Tensor, device=None, is_in_use=False):
        self.chunk_shape, self.chunk_dtype = chunk.shape, chunk.dtype
        self.device = device or chunk.device
        if torch.cuda.is_available() and chunk.is_cuda:
            host = torch.empty(chunk.shape, dtype=chunk.dtype, device="cpu", pin_memory=True)
            host.copy_(chunk, non_blocking=True)
            self.cpu_chunk = host
        else:
            self.cpu_chunk = chunk.detach().cpu()
        self.gpu_chunk = chunk if is_in_use else None

    def load_to_gpu(self):
        if self.gpu_chunk is None:
            self.gpu_chunk = self.cpu_chunk.to(self.device, non_blocking=True)

    def get_gpu_chunk(self):
        assert self.gpu_chunk is not None and self.gpu_chunk.device == self.device
        return self.gpu_chunk

    def check_gpu_chunk(self):
        return self.gpu_chunk is not None

    def offload(self):
        self.gpu_chunk = None

    def overwrite_to_cpu(self):
        assert self.gpu_chunk is not None
        self.cpu_chunk.copy_(self.gpu_chunk, non_blocking=True)


def FPDT_InputConstruct(tokens, labels, loss_mask, attention_mask, position_ids, args=None, sp_size=1, sp_rank=0,
                        num_chunks=1):
    """Load-balanced chunk assignment: the global sequence is cut into ``sp_size * num_chunks`` pieces and rank
    ``r`` takes pieces ``r, r + sp_size, ...`` so every rank owns early *and* late (cheap and expensive causal)
    positions (reference :79)."""
    seq = tokens.shape[1]
    assert seq % (sp_size * num_chunks) == 0
    piece = seq // (sp_size * num_chunks)
    idx = torch.cat([torch.arange((c * sp_size + sp_rank) * piece, (c * sp_size + sp_rank + 1) * piece)
                     for c in range(num_chunks)]).to(tokens.device)
    take = lambda t: None if t is None else t.index_select(1, idx)
    return take(tokens), take(labels), take(loss_mask), attention_mask, take(position_ids)


class FPDT_Attention(torch.nn.Module):
    """Chunked causal self-attention with fused QKV / output projections (reference :971).

    ``forward(hidden_states [S_local, B, H])`` -> ``[S_local, B, H]``.  ``qkv_linear_weight`` is
    ``[3*H_proj, H]`` laid out ``[q | k | v]``; ``num_heads`` may differ from ``num_kv_heads`` (GQA)."""

    def __init__(self, config=None, first_weight=None, first_bias=None, second_weight=None, second_bias=None,
                 sequence_process_group=None, gather_idx: int = 0, scatter_idx: int = 2, return_bias=True,
                 chunk_size=65536, enable_offloading=True, num_heads=None, num_kv_heads=None, head_dim=None):
        super().__init__()
        self.qkv_linear_weight, self.qkv_linear_bias = first_weight, first_bias
        self.qkv_dense_weight, self.qkv_dense_bias = second_weight, second_bias
        self.spg = sequence_process_group
        self.chunk_size = chunk_size
        self.enable_offloading = enable_offloading
        self.return_bias = return_bias
        self.num_heads = num_heads or getattr(config, "num_attention_heads")
        self.num_kv_heads = num_kv_heads or getattr(config, "num_key_value_heads", self.num_heads)
        hidden = getattr(config, "hidden_size", None) or first_weight.shape[1]
        self.head_dim = head_dim or hidden // self.num_heads

    def _proj(self, x):
        return F.linear(x, self.qkv_linear_weight, self.qkv_linear_bias)

    def forward(self, hidden_states, attention_mask=None, rotary_pos_emb=None, cpu_offloading=None):
        S, B, _ = hidden_states.shape
        sp = dist.get_world_size(self.spg) if self.spg is not None else 1
        offload = self.enable_offloading if cpu_offloading is None else cpu_offloading
        offload = offload and hidden_states.is_cuda
        hq, hkv, d = self.num_heads, self.num_kv_heads, self.head_dim
        n_chunks = max(1, math.ceil(S * sp / self.chunk_size))
        while S % n_chunks:
            n_chunks += 1
        cs = S // n_chunks
        scale = 1.0 / math.sqrt(d)
        k_chunks, v_chunks = [], []
        outs = []
        from torch.utils.checkpoint import checkpoint
        for i in range(n_chunks):
            x = hidden_states[i * cs:(i + 1) * cs]
            qkv = self._proj(x)  # [cs, B, (hq+2hkv)*d]
            q, k, v = torch.split(qkv, [hq * d, hkv * d, hkv * d], dim=-1)
            q = q.reshape(cs, B, hq, d)
            k = k.reshape(cs, B, hkv, d)
            v = v.reshape(cs, B, hkv, d)
            if sp > 1:  # Ulysses: sequence gather / head scatter
                q = _SeqAllToAll.apply(self.spg, q, 2, 0, 1)
                k = _SeqAllToAll.apply(self.spg, k, 2, 0, 1)
                v = _SeqAllToAll.apply(self.spg, v, 2, 0, 1)
            q, k, v = (t.permute(1, 2, 0, 3) for t in (q, k, v))  # [B, h, s, d]
            k_chunks.append(SequenceChunk(k, is_in_use=True) if offload else k)
            v_chunks.append(SequenceChunk(v, is_in_use=True) if offload else v)
            out = lse = None
            for j in range(i + 1):
                if offload:
                    if j + 1 <= i:  # prefetch the next pair while this one computes
                        k_chunks[j + 1].load_to_gpu()
                        v_chunks[j + 1].load_to_gpu()
                    k_chunks[j].load_to_gpu()
                    v_chunks[j].load_to_gpu()
                    kj, vj = k_chunks[j].get_gpu_chunk(), v_chunks[j].get_gpu_chunk()
                else:
                    kj, vj = k_chunks[j], v_chunks[j]
                if torch.is_grad_enabled() and q.requires_grad:
                    bo, bl = checkpoint(_attn_with_lse, q, kj, vj, j == i, scale, use_reentrant=False)
                else:
                    bo, bl = _attn_with_lse(q, kj, vj, j == i, scale)
                out, lse = update_out_and_lse(out, lse, bo, bl)
                if offload and j < i:
                    k_chunks[j].offload()
                    v_chunks[j].offload()
            o = out.to(hidden_states.dtype).permute(2, 0, 1, 3)  # [s_full_chunk, B, h_local, d]
            if sp > 1:
                o = _SeqAllToAll.apply(self.spg, o, 0, 2, 1)
            outs.append(o.reshape(cs, B, hq * d))
            if offload:
                k_chunks[i].offload()
                v_chunks[i].offload()
        ctx = torch.cat(outs, dim=0)
        y = F.linear(ctx, self.qkv_dense_weight, None if self.return_bias else self.qkv_dense_bias)
        if self.return_bias:
            return y, self.qkv_dense_bias
        return y


class FPDT_FFN(torch.nn.Module):
    """MLP evaluated chunk by chunk under activation checkpointing so only one chunk's intermediate
    ``[chunk, B, 4H]`` activation is alive (reference :1056)."""

    def __init__(self, first_weight=None, first_bias=None, second_weight=None, second_bias=None, chunk_size=65536,
                 activation=F.gelu, add_bias=True):
        super().__init__()
        self.w1, self.b1, self.w2, self.b2 = first_weight, first_bias, second_weight, second_bias
        self.chunk_size = chunk_size
        self.activation = activation
        self.add_bias = add_bias

    def _chunk(self, x):
        h = self.activation(F.linear(x, self.w1, self.b1))
        return F.linear(h, self.w2, self.b2 if self.add_bias else None)

    def forward(self, x):
        from torch.utils.checkpoint import checkpoint
        S = x.shape[0]
        outs = []
        for s in range(0, S, self.chunk_size):
            xc = x[s:s + self.chunk_size]
            outs.append(checkpoint(self._chunk, xc, use_reentrant=False) if (torch.is_grad_enabled() and
                                                                             x.requires_grad) else self._chunk(xc))
        y = torch.cat(outs, dim=0)
        return y if self.add_bias else (y, self.b2)


class FPDT_LogitsLoss(torch.nn.Module):
    """Vocabulary projection + cross entropy over token chunks: the ``[tokens, vocab]`` logits never
    materialise (reference :1137).  Backed by the framework's fused chunked kernel path."""

    def __init__(self, lm_head_weight, chunk_size=2048, ignore_index=-100):
        super().__init__()
        self.weight = lm_head_weight
        self.chunk_size = chunk_size
        self.ignore_index = ignore_index

    def forward(self, hidden, labels):
        from deepspeed_b200.ops.linear import chunked_linear_xent
        return chunked_linear_xent(hidden.reshape(-1, hidden.shape[-1]), self.weight, labels.reshape(-1),
                                   chunk=self.chunk_size, ignore_index=self.ignore_index)


# ---- element-wise helpers of the reference module (``fpdt_layer.py:32, :1044-1053``) -----------------------------------
def _rotate_half_backward(x):
    x1, x2 = x.chunk(2, dim=-1)
    return torch.cat((x2, -x1), dim=-1)


def apply_rotary_pos_emb_backward(grad_output, freqs_cos, freqs_sin):
    """Adjoint of ``sequence.layer.apply_rotary_pos_emb``."""
    rot = freqs_cos.shape[-1]
    g, g_pass = grad_output[..., :rot], grad_output[..., rot:]
    g = g * freqs_cos + _rotate_half_backward(g * freqs_sin)
    return g if g_pass.shape[-1] == 0 else torch.cat((g, g_pass), dim=-1)


_GELU_K, _GELU_C = 0.7978845608028654, 0.044715


def bias_gelu(x):
    """tanh-approximated GELU."""
    return 0.5 * x * (1.0 + torch.tanh(_GELU_K * x * (1.0 + _GELU_C * x * x)))


def bias_gelu_back(g, x):
    th = torch.tanh(_GELU_K * x * (1.0 + _GELU_C * x * x))
    d = 0.5 * (1.0 + th) + 0.5 * x * (1.0 - th * th) * _GELU_K * (1.0 + 3.0 * _GELU_C * x * x)
    return d * g
