"""ZeroQuant-style quantizer objects, packed-parameter helpers, data-pipeline / tuner utilities."""
import numpy as np
import pytest
import torch


@pytest.mark.parametrize("bits", [8, 4])
@pytest.mark.parametrize("group_dim", [0, 1])
def test_groupwise_asymmetric_roundtrip(bits, group_dim):
    from deepspeed_b200.inference.quantization.utils import DeQuantizer, Quantizer
    torch.manual_seed(0)
    w = torch.randn(64, 96)
    conf = {"num_bits": bits, "group_size": 32, "group_dim": group_dim, "symmetric": False}
    codes, scale, mn = Quantizer(conf).quantize(w)
    assert codes.dtype == torch.uint8 and codes.shape == ((64, 96) if bits == 8 else (64, 48))
    back = DeQuantizer(conf, torch.float32).dequantize(codes, scale, mn)
    # error bounded by half a quantization step of each group
    g = w.reshape(2, 32, 96) if group_dim == 0 else w.reshape(64, 3, 32)
    ax = 1 if group_dim == 0 else 2
    step = (g.amax(ax, keepdim=True) - g.amin(ax, keepdim=True)) / (2**bits - 1)
    err = (back - w).reshape(g.shape).abs()
    assert (err <= step * 0.5 + 1e-5).all()


def test_param_packing_and_functional_wrap():
    from deepspeed_b200.inference.quantization.utils import (_quantize_param, dequantize_param, recursive_setattr,
                                                            wrap_quantized_functional)
    torch.manual_seed(0)
    lin = torch.nn.Linear(64, 32, bias=False)
    ref = lin.weight.detach().clone()
    _quantize_param(lin.weight, {"num_bits": 8, "group_size": 16, "group_dim": 1, "symmetric": False})
    assert lin.weight.weight_quantized and lin.weight.dtype == torch.uint8 and lin.weight.dim() == 1
    assert lin.weight.numel() == 32 * 64 + 2 * (32 * 4) * 4  # codes + fp32 scale + fp32 min per group
    deq = dequantize_param(lin.weight)
    assert deq.shape == ref.shape and (deq - ref).abs().max() < 0.02
    x = torch.randn(3, 64)
    y = wrap_quantized_functional(torch.nn.functional.linear)(x, lin.weight)
    assert torch.allclose(y, x @ deq.t(), atol=1e-5)
    with pytest.raises(AssertionError):
        _quantize_param(lin.weight, {"num_bits": 8, "group_size": 16, "group_dim": 1, "symmetric": False})
    m = torch.nn.Sequential(torch.nn.Sequential(torch.nn.Linear(2, 2)))
    recursive_setattr(m, "0.0", torch.nn.Identity())
    assert isinstance(m[0][0], torch.nn.Identity)


def test_data_pipeline_and_tuner_utils(tmp_path):
    from deepspeed_b200.autotuning.tuner.utils import (dict_to_dims, dict_to_feature, feature_to_index, flatten,
                                                      gen_combinations, index_to_feature)
    from deepspeed_b200.runtime.data_pipeline.data_routing.utils import (bsh_decoder_gather, bsh_decoder_scatter,
                                                                        sbh_decoder_gather, sbh_decoder_scatter)
    from deepspeed_b200.runtime.data_pipeline.data_sampling.utils import (close_mmap_dataset_builder,
                                                                         create_mmap_dataset_builder, find_fit_int_dtype,
                                                                         split_dataset, split_index)
    assert find_fit_int_dtype(0, 255) is np.uint8 and find_fit_int_dtype(0, 70000) is np.uint32
    assert find_fit_int_dtype(-5, 100) is np.int8 and find_fit_int_dtype(-40000, 5) is np.int32
    assert [tuple(map(int, s)) for s in split_index(0, 10, 3)] == [(0, 3), (3, 6), (6, 10)]
    ws, ts = split_dataset(list(range(100)), 4, 1, 2)
    assert tuple(map(int, ws[1])) == (25, 50) and [tuple(map(int, t)) for t in ts] == [(25, 37), (37, 50)]
    b = create_mmap_dataset_builder(str(tmp_path / "m"), np.int32)
    b.add_item(torch.tensor([1, 2, 3]))
    close_mmap_dataset_builder(b, str(tmp_path / "m"))
    assert (tmp_path / "m.idx").exists() and (tmp_path / "m.bin").exists()
    h = torch.arange(2 * 6 * 3, dtype=torch.float32).reshape(2, 6, 3)
    mask = torch.ones(2, 1, 6, 6)
    part, idx, pm = bsh_decoder_gather(4, h, mask)
    assert part.shape == (2, 4, 3) and pm.shape == (2, 1, 4, 4) and all((i[1:] > i[:-1]).all() for i in idx)
    out = bsh_decoder_scatter(torch.zeros_like(h), part, idx)
    assert torch.equal(out[0, idx[0]], h[0, idx[0]])
    hs = h.transpose(0, 1).contiguous()
    part, idx, _ = sbh_decoder_gather(4, hs, mask)
    assert part.shape == (4, 2, 3)
    out = sbh_decoder_scatter(torch.zeros_like(hs), part, idx)
    assert torch.equal(out[idx[1], 1], hs[idx[1], 1])
    space = {"a": [1, 2, 3], "z": {"s": [0, 1], "b": 5}}
    dims = dict_to_dims(space)
    assert dims == [3, 2, 1]
    for p in range(6):
        assert feature_to_index(index_to_feature(p, dims), dims) == p
    assert len(list(gen_combinations(space))) == 6
    assert flatten({"a": {"b": 1, "c": {"d": 2}}}) == {"a_b": 1, "a_c_d": 2}
    assert dict_to_feature({"x": 2, "y": "auto", "n": {"x": 4}}, ["x", "n"]) == [2.0, 4.0]


def _qctx():
    import torch.distributed as td
    from deepspeed_b200.inference.quantization.quantization_context import QuantizationContext
    from deepspeed_b200.inference.quantization.utils import dequantize_param
    from deepspeed_b200.runtime.zero.partition_parameters import GatheredParameters, is_zero_param
    cfg = {"train_micro_batch_size_per_gpu": 1, "zero_optimization": {"stage": 3},
           "weight_quantization": {"post_init_quant": {"Linear.weight": {"num_bits": 8, "group_size": 16, "group_dim": 1}}}}
    torch.manual_seed(0)
    ref = torch.nn.Linear(64, 32)
    torch.manual_seed(0)
    with QuantizationContext(cfg):
        m = torch.nn.Sequential(torch.nn.Linear(64, 32))
    w = m[0].weight
    assert is_zero_param(w) and w.weight_quantized and w.quant_full_shape == (32, 64)
    assert not getattr(m[0].bias, "weight_quantized", False)
    with GatheredParameters([w]):
        assert w.dtype == torch.uint8
        deq = dequantize_param(w)
    assert (deq - ref.weight).abs().max() < 0.01


def test_quantization_context_shards_packed_weights():
    from tests.common import run_distributed
    run_distributed(_qctx, 2)


@pytest.mark.parametrize("mode,tol", [("int8", 0.02), ("int4", 0.2), ("fp8", 0.08), ("fp6", 0.2)])
def test_weight_only_quantized_linear_all_formats_host(mode, tol):
    from deepspeed_b200.inference.quantization.layers import maybe_quantized_linear, quantize_weight
    torch.manual_seed(0)
    w = torch.randn(32, 128) * 0.1
    x = torch.randn(3, 128)
    qw = quantize_weight(w, mode, group_size=64)
    deq = qw.dequantize()
    assert deq.shape == w.shape and (deq - w).abs().max() < tol * w.abs().max()
    y = maybe_quantized_linear(x, qw)
    assert torch.allclose(y, x @ deq.t(), atol=1e-5)


def test_quantized_embedding_and_post_init_quant():
    import torch
    from deepspeed_b200.inference.quantization import _init_group_wise_weight_quantization
    from deepspeed_b200.inference.quantization.layers import QuantizedEmbedding, QuantizedLinear

    class M(torch.nn.Module):

        def __init__(self):
            super().__init__()
            self.embed = torch.nn.Embedding(32, 128)
            self.proj = torch.nn.Linear(128, 128)

    m = M()
    ref = m.embed(torch.arange(32))
    _init_group_wise_weight_quantization(m, {"weight_quantization": {"post_init_quant": {"embed": {"num_bits": 8, "group_size": 64},
                                                                                      "proj": {"num_bits": 8, "group_size": 64}}}})
    assert isinstance(m.embed, QuantizedEmbedding) and isinstance(m.proj, QuantizedLinear)
    assert m.embed.weight.dtype == torch.uint8 and (m.embed(torch.arange(32)) - ref).abs().max() < 0.05
