"""Checkpoint-time int8 weight quantisation for the v1 inference path (reference ``runtime/weight_quantizer.py:10``).

Each quantised tensor is cut into ``groups`` equal pieces of its flattened data; a piece is stored as
``round(x * s)`` with ``s = 2^bits / (2 * max|x| + 1e-5)`` and the *inverse* scales are collected per layer in the order
[qkv, attention-out, mlp-in, mlp-out] (rows zero-padded to the widest), which is the layout the fused layer's
dequantising GEMMs index.
"""
import torch

_KINDS = ("qkv", "dense", "h4h", "4hh")
_KEY_OF = {"attention.query_key_value.weight": "qkv", "mlp.dense_h_to_4h.weight": "h4h", "mlp.dense_4h_to_h.weight": "4hh",
           "attention.dense.weight": "dense"}


def _inverse_row(scale):
    return (1.0 / scale).reshape(1, -1)


def _pad_rows(rows):
    width = max(r.shape[-1] for r in rows)
    return torch.cat([torch.nn.functional.pad(r, (0, width - r.shape[-1])) for r in rows], dim=0)


class WeightQuantization:

    def __init__(self, mlp_extra_grouping=True, mp_size=1):
        self.mlp_extra_grouping = mlp_extra_grouping
        self.mp_size = mp_size
        self._scales = {k: [] for k in _KINDS}  # per kind: one [1, groups] inverse-scale row per layer

    # reference attribute names
    dense_scales = property(lambda self: self._scales["dense"])
    qkv_scales = property(lambda self: self._scales["qkv"])
    mlp4hh_scales = property(lambda self: self._scales["4hh"])
    mlph4h_scales = property(lambda self: self._scales["h4h"])

    # ---- tensor level -------------------------------------------------------------------------------------------------
    def quantize_data(self, data, quantize_bits, groups, key=None):
        """-> (int8 tensor shaped like ``data``, scales [groups, 1])."""
        g = data.detach().float().reshape(groups, -1)
        levels = 2**quantize_bits
        bound = torch.maximum(g.amax(dim=1), g.amin(dim=1).abs())
        scale = levels / (2 * bound + 1e-5)
        q = (g * scale[:, None]).round_().clamp_(-levels // 2, levels // 2 - 1)
        return q.reshape(data.shape).to(torch.int8), scale.reshape(groups, 1)

    def _ratio(self, data, merge_count=1):
        a, b = self.mp_size * data.shape[0] * merge_count, self.mp_size * data.shape[1] * merge_count
        return a / data.shape[1], b / data.shape[0]

    def is_mlp(self, data, merge_count=1):
        return 4 in self._ratio(data, merge_count)

    def is_qkv(self, data):
        return 3 in self._ratio(data)

    def Quantize(self, value_list, quantize_bits, groups, key, merge_dim=0):
        """Quantise every shard in ``value_list`` (in place) and record the layer's inverse scales under ``key``'s kind."""
        if self.mlp_extra_grouping and self.is_mlp(value_list[0], merge_count=len(value_list)):
            groups *= 2
        scales = []
        for i, shard in enumerate(value_list):
            value_list[i], s = self.quantize_data(shard, quantize_bits, groups, key)
            scales.append(s)
        row = _inverse_row(torch.cat(scales, dim=merge_dim)).to(value_list[0].device)
        kind = next((k for pat, k in _KEY_OF.items() if pat in key and k != "dense"), "dense")
        self._scales[kind].append(row)
        return value_list

    # ---- scale assembly -------------------------------------------------------------------------------------------------
    def merge_layer_scales(self, layer_scales):
        return _pad_rows(list(layer_scales)).unsqueeze(0)

    def _layers(self):
        s = self._scales
        return zip(s["qkv"], s["dense"], s["h4h"], s["4hh"])

    def merge_scales(self):
        return torch.cat([self.merge_layer_scales(layer) for layer in self._layers()])

    def merge_scales_split(self, split_count):
        """Scales of a checkpoint being re-split over ``split_count`` ranks: every kind's row is cut evenly, qkv and dense
        rows are doubled in width (zero filled) to line up with the MLP rows."""
        per_rank = [[] for _ in range(split_count)]
        for qkv, dense, h4h, o4h in self._layers():
            cut = [torch.split(t, t.numel() // split_count, dim=-1) for t in (qkv, dense, h4h, o4h)]
            for r in range(split_count):
                q, d, a, b = (c[r] for c in cut)
                z = torch.zeros_like(q)
                per_rank[r].append(torch.cat([torch.cat((q, z), dim=1), torch.cat((d, z), dim=1), a, b]).unsqueeze(0))
        return [torch.cat(rows) for rows in per_rank]

    # ---- whole state dict / model ---------------------------------------------------------------------------------------
    def sd_quantize_megatron(self, sd, quantize_bits, groups):
        for key in list(sd.keys()):
            if any(pat in key for pat in _KEY_OF):
                sd[key] = self.Quantize([sd[key]], quantize_bits, groups, key=key)[0]
        return sd, self.merge_scales()

    def model_quantize(self, model, quantize_policy, quantize_bits, groups):
        """Quantise the four GEMM weights of every layer whose class has a policy in ``quantize_policy``."""
        collected = []

        def visit(module):
            for name, child in module.named_children():
                policy_cls = quantize_policy.get(child.__class__)
                if policy_cls is None:
                    visit(child)
                    continue
                policy = policy_cls(child)
                att, mlp = policy.attention(), policy.mlp()
                weights = [att[0] if torch.is_tensor(att[0]) and att[0].dim() == 2 else att[1], att[2] if len(att) == 4 else att[3],
                           mlp[0] if torch.is_tensor(mlp[0]) and mlp[0].dim() == 2 else mlp[1], mlp[2] if len(mlp) == 4 else mlp[3]]
                rows = []
                for w in weights:
                    q, s = self.quantize_data(w.data, quantize_bits, groups)
                    w.data = q
                    rows.append(_inverse_row(s).to(w.device))
                collected.append(self.merge_layer_scales(rows))

        visit(model)
        return model, torch.cat(collected)
