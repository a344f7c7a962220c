import copy

import torch
from torch import nn

from deepspeed_b200.linear import (Init, LoRAConfig, LoRAOptimizedLinear, OptimizedLinear, QuantizationConfig, QuantizedLinear,
                                   QuantizedParameter)


def test_factory_and_lora_training():
    assert isinstance(OptimizedLinear(8, 4, dtype=torch.float32), nn.Linear)
    torch.manual_seed(0)
    l = OptimizedLinear(32, 16, lora_config=LoRAConfig(lora_r=4, lora_alpha=8), dtype=torch.float32)
    assert isinstance(l, LoRAOptimizedLinear) and not l.weight.requires_grad
    x = torch.randn(5, 32)
    base = x @ l.weight.t()
    torch.testing.assert_close(l(x), base)          # B = 0 -> adapter is the identity at init
    opt = torch.optim.SGD([p for p in l.parameters() if p.requires_grad], lr=0.1)
    tgt = torch.randn(5, 16)
    l0 = None
    for _ in range(20):
        loss = (l(x) - tgt).pow(2).mean()
        l0 = l0 or loss.item()
        opt.zero_grad()
        loss.backward()
        opt.step()
    assert loss.item() < l0 and l.weight.grad is None
    sd = {"weight": torch.randn(16, 32)}
    l.load_state_dict(sd, strict=False)
    torch.testing.assert_close(l.weight.data, sd["weight"])


def test_quantized_parameter_and_linear():
    torch.manual_seed(0)
    w = torch.randn(64, 128, dtype=torch.bfloat16)
    qp = QuantizedParameter(w.clone(), quantization_config=QuantizationConfig(q_bits=8, group_size=128))
    assert qp._scale is not None and qp._scale.numel() == w.numel() // 128
    dq = qp.dequantized()
    assert dq.shape == w.shape and (dq.float() - w.float()).abs().max() < 0.3
    ql = QuantizedLinear(128, 64, quantization_config=QuantizationConfig(q_bits=8, group_size=128))
    y = ql(torch.randn(3, 128, dtype=torch.bfloat16))
    assert y.shape == (3, 64)
    q2 = copy.deepcopy(qp)
    torch.testing.assert_close(q2.dequantized(), dq)
    lq = OptimizedLinear(128, 64, lora_config=LoRAConfig(lora_r=4), quantization_config=QuantizationConfig(group_size=128))
    assert isinstance(lq.weight, QuantizedParameter)
    assert lq(torch.randn(2, 128, dtype=torch.bfloat16)).shape == (2, 64)


def test_init_context_swaps_linears():
    with Init(lora_config=LoRAConfig(lora_r=2), quant_config=None):
        m = nn.Sequential(nn.Linear(8, 8, bias=False, dtype=torch.float32), nn.Linear(8, 4, bias=True))
    assert isinstance(m[0], LoRAOptimizedLinear) and type(m[1]) is nn.Linear
    assert type(nn.Linear(2, 2)) is nn.Linear
