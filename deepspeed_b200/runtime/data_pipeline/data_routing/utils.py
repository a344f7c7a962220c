"""Pure-torch random-LTD token gather / scatter for decoder layouts (reference
``runtime/data_pipeline/data_routing/utils.py``); the kernel path is ``ops/random_ltd``."""
import torch


def _pick(n_tokens, keep, device):
    return torch.randperm(n_tokens, device=device)[:keep].sort()[0]


def bsh_decoder_gather(reserved_length, hidden_states, mask):
    """[batch, seq, hidden]: keep ``reserved_length`` random (order-preserving) tokens per sample."""
    idx = [_pick(hidden_states.size(1), reserved_length, hidden_states.device) for _ in range(hidden_states.size(0))]
    part = torch.stack([hidden_states[b, i] for b, i in enumerate(idx)], dim=0)
    return part, idx, mask[:, :, :reserved_length, :reserved_length]


def bsh_decoder_scatter(hidden_states, part_hidden_states, rand_list):
    for b, i in enumerate(rand_list):
        hidden_states[b, i, :] = part_hidden_states[b]
    return hidden_states


def sbh_decoder_gather(reserved_length, hidden_states, mask):
    """[seq, batch, hidden] variant (Megatron layout)."""
    idx = [_pick(hidden_states.size(0), reserved_length, hidden_states.device) for _ in range(hidden_states.size(1))]
    part = torch.stack([hidden_states[i, b] for b, i in enumerate(idx)], dim=1)
    return part, idx, mask[:, :, :reserved_length, :reserved_length]


def sbh_decoder_scatter(hidden_states, part_hidden_states, rand_list):
    for b, i in enumerate(rand_list):
        hidden_states[i, b, :] = part_hidden_states[:, b]
    return hidden_states
