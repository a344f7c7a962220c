import copy

import torch
from torch import nn

from deepspeed_b200.compression import compression_scheduler, init_compression, redundancy_clean
from deepspeed_b200.compression.basic_layer import LinearLayer_Compress
from deepspeed_b200.compression.config import get_compression_config


class Block(nn.Module):

    def __init__(self):
        super().__init__()
        self.qkv = nn.Linear(16, 48)
        self.out = nn.Linear(16, 16)
        self.fc1 = nn.Linear(16, 32)
        self.fc2 = nn.Linear(32, 16)

    def forward(self, x):
        q, k, v = self.qkv(x).chunk(3, -1)
        a = torch.softmax(q @ k.transpose(-1, -2) / 4.0, -1) @ v
        x = x + self.out(a)
        return x + self.fc2(torch.relu(self.fc1(x)))


class Net(nn.Module):

    def __init__(self, n=2):
        super().__init__()
        self.layers = nn.ModuleList([Block() for _ in range(n)])

    def forward(self, x):
        for l in self.layers:
            x = l(x)
        return x


CFG = {"compression_training": {
    "weight_quantization": {"shared_parameters": {"enabled": True, "quantize_weight_in_forward": True, "quantize_groups": 4,
                                                  "schedule_offset": 2},
                            "different_groups": {"wq": {"params": {"start_bits": 8, "target_bits": 8}, "modules": ["fc"]}}},
    "activation_quantization": {"shared_parameters": {"enabled": True, "schedule_offset": 0, "range_calibration": "dynamic"},
                                "different_groups": {"aq": {"params": {"bits": 8}, "modules": ["fc1"]}}},
    "sparse_pruning": {"shared_parameters": {"enabled": True, "schedule_offset": 0, "method": "l1"},
                       "different_groups": {"sp": {"params": {"dense_ratio": 0.5}, "modules": ["out"]}}},
    "row_pruning": {"shared_parameters": {"enabled": True, "schedule_offset": 0, "method": "l1"},
                    "different_groups": {"rp": {"params": {"dense_ratio": 0.5}, "modules": ["fc1"],
                                                "related_modules": [["fc2"]]}}},
}}


def test_config_defaults():
    c = get_compression_config(CFG)
    assert c["weight_quantization"]["shared_parameters"]["quantization_type"] == "symmetric"
    assert c["head_pruning"]["shared_parameters"]["enabled"] is False
    assert c["row_pruning"]["different_groups"]["rp"]["related_modules"] == [["fc2"]]


def test_init_schedule_and_clean():
    torch.manual_seed(0)
    model = Net()
    dense = copy.deepcopy(model)
    init_compression(model, CFG)
    assert isinstance(model.layers[0].fc1, LinearLayer_Compress)
    sched = compression_scheduler(model, get_compression_config(CFG))
    x = torch.randn(3, 5, 16)
    sched.step(step_zero_check=True)
    fc1 = model.layers[0].fc1
    assert fc1.row_pruning_enabled and fc1.activation_quantization_enabled and not fc1.weight_quantization_enabled
    y0 = model(x)
    assert not torch.allclose(y0, dense(x))
    sched.step(), sched.step()
    assert fc1.weight_quantization_enabled
    y1 = model(x)
    y1.sum().backward()          # straight-through gradients reach the dense weights
    assert fc1.weight.grad is not None and torch.isfinite(fc1.weight.grad).all()
    # sparse mask: half of `out` weights are zeroed in the effective weight
    out = model.layers[0].out
    assert abs(out.get_mask("sparse").float().mean().item() - 0.5) < 0.05
    redundancy_clean(model, CFG)
    assert model.layers[0].fc1.out_features == 16 and model.layers[0].fc2.in_features == 16   # rows physically removed
    assert (model.layers[0].out.weight == 0).float().mean() > 0.45
    y2 = model.eval()(x)
    assert y2.shape == x.shape and torch.isfinite(y2).all()


def test_layer_reduction():
    from deepspeed_b200.compression import student_initialization
    teacher, student = Net(4), Net(2)
    cfg = {"compression_training": {"layer_reduction": {"enabled": True, "keep_number_layer": 2,
                                                         "module_name_prefix": "layers", "teacher_layer": [1, 3],
                                                         "other_module_name": []}}}
    student_initialization(student, teacher, cfg)
    assert torch.equal(student.layers[1].fc1.weight, teacher.layers[3].fc1.weight)
