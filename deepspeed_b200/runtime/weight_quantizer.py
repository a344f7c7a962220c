"""Checkpoint-time weight quantisation for inference (reference ``runtime/weight_quantizer.py:10``)."""
import torch


class WeightQuantization:

    def __init__(self, mlp_extra_grouping=True, mp_size=1):
        self.dense_scales, self.qkv_scales, self.mlp4hh_scales, self.mlph4h_scales = [], [], [], []
        self.mlp_extra_grouping = mlp_extra_grouping
        self.mp_size = mp_size

    def quantize_data(self, data, quantize_bits, groups, key=None):
        chunks = torch.split(data.float().flatten(), data.numel() // groups)
        q_range = 2**quantize_bits
        scales = [q_range / (2 * max(c.max(), c.min().abs()) + 1e-5) for c in chunks]
        q = [(c * s).round().clamp(-q_range // 2, q_range // 2 - 1) for c, s in zip(chunks, scales)]
        data_int = torch.cat(q).reshape(data.shape).to(torch.int8)
        scale = torch.cat([s.unsqueeze(0).unsqueeze(0) for s in scales])
        return data_int, scale

    def is_mlp(self, data, merge_count=1):
        return (self.mp_size * data.shape[0] * merge_count) / data.shape[1] == 4 or \
            (self.mp_size * data.shape[1] * merge_count) / data.shape[0] == 4

    def is_qkv(self, data):
        return (self.mp_size * data.shape[0]) / data.shape[1] == 3 or (self.mp_size * data.shape[1]) / data.shape[0] == 3

    def Quantize(self, value_list, quantize_bits, groups, key, merge_dim=0):
        if self.mlp_extra_grouping and self.is_mlp(value_list[0], merge_count=len(value_list)):
            groups *= 2
        q_scale = []
        for i, data in enumerate(value_list):
            data_int, scale = self.quantize_data(data, quantize_bits, groups, key)
            q_scale.append(scale)
            value_list[i] = data_int
        q_scale = 1 / torch.cat(q_scale, dim=merge_dim).to(value_list[0].device if value_list[0].is_cuda else "cpu") \
            .view(-1).unsqueeze(0)
        if "mlp.dense_4h_to_h.weight" in key:
            self.mlp4hh_scales.append(q_scale)
        elif "mlp.dense_h_to_4h.weight" in key:
            self.mlph4h_scales.append(q_scale)
        elif "attention.query_key_value.weight" in key:
            self.qkv_scales.append(q_scale)
        else:
            self.dense_scales.append(q_scale)
        return value_list

    def merge_layer_scales(self, layer_scales):
        max_dim = max(s.shape[-1] for s in layer_scales)
        layer_scales = [torch.cat((s, torch.zeros((1, max_dim - s.shape[-1]), device=s.device)), dim=-1)
                        if s.shape[-1] < max_dim else s for s in layer_scales]
        return torch.cat(layer_scales).unsqueeze(0)

    def merge_scales(self):
        all_scales = []
        for dense, qkv, m4hh, mh4h in zip(self.dense_scales, self.qkv_scales, self.mlp4hh_scales, self.mlph4h_scales):
            all_scales.append(self.merge_layer_scales([qkv, dense, mh4h, m4hh]))
        return torch.cat(all_scales)

    def merge_scales_split(self, split_count):
        all_scales = [[] for _ in range(split_count)]
        for dense, qkv, m4hh, mh4h in zip(self.dense_scales, self.qkv_scales, self.mlp4hh_scales, self.mlph4h_scales):
            d = torch.split(dense, dense.numel() // split_count)
            q = torch.split(qkv, qkv.numel() // split_count)
            a = torch.split(m4hh, m4hh.numel() // split_count)
            b = torch.split(mh4h, mh4h.numel() // split_count)
            for s in range(split_count):
                all_scales[s].append(torch.cat([torch.cat((q[s], torch.zeros_like(q[s])), dim=1),
                                                torch.cat((d[s], torch.zeros_like(q[s])), dim=1), b[s], a[s]]).unsqueeze(0))
        return [torch.cat(s) for s in all_scales]

    def sd_quantize_megatron(self, sd, quantize_bits, groups):
        keys = sd.keys()
        for key in keys:
            value_list = [sd[key]]
            if any(k in key for k in ("attention.dense.weight", "mlp.dense_4h_to_h.weight", "mlp.dense_h_to_4h.weight",
                                      "attention.query_key_value.weight")):
                value_list = self.Quantize(value_list, quantize_bits, groups, key=key)
            sd[key] = value_list[0]
        return sd, self.merge_scales()

    def model_quantize(self, model, quantize_policy, quantize_bits, groups):
        all_scales = []

        def quantize_fn(layer, policy_cls):
            policy = policy_cls(layer)
            _, qkvw, _, dense_w, _, _ = policy.attention()
            _, _h4h_w, _, _4hh_w, _ = policy.mlp()
            keys = [qkvw, dense_w, _h4h_w, _4hh_w]
            layer_scales = []
            for k in keys:
                q, s = self.quantize_data(k.data, quantize_bits, groups)
                k.data = q
                layer_scales.append(1 / s.to(k.device).view(-1).unsqueeze(0))
            all_scales.append(self.merge_layer_scales(layer_scales))
            return layer

        def walk(m):
            for name, child in m.named_children():
                if child.__class__ in quantize_policy:
                    setattr(m, name, quantize_fn(child, quantize_policy[child.__class__]))
                else:
                    walk(child)
            return m

        return walk(model), torch.cat(all_scales)
