import torch
from torch import nn

from deepspeed_b200.profiling.flops_profiler import FlopsProfiler, get_model_profile


class Net(nn.Module):

    def __init__(self):
        super().__init__()
        self.a = nn.Linear(16, 32)
        self.b = nn.Sequential(nn.ReLU(), nn.Linear(32, 8, bias=False))

    def forward(self, x):
        return self.b(self.a(x))


def test_counts_and_tree(tmp_path):
    m = Net()
    flops, macs, params = get_model_profile(m, input_shape=(4, 16), print_profile=True, as_string=False,
                                            output_file=str(tmp_path / "p.txt"), module_depth=-1, top_modules=2)
    assert macs == 4 * 16 * 32 + 4 * 32 * 8
    assert params == 16 * 32 + 32 + 32 * 8
    assert flops >= 2 * macs
    txt = (tmp_path / "p.txt").read_text()
    assert "Aggregated Profile" in txt and "Linear" in txt and "fwd MACs per GPU" in txt


def test_start_stop_api_and_conv_sdpa():
    conv = nn.Conv2d(3, 4, 3, padding=1, bias=False)
    prof = FlopsProfiler(conv)
    prof.start_profile()
    conv(torch.randn(2, 3, 8, 8))
    prof.stop_profile()
    assert prof.get_total_macs() == 2 * 4 * 8 * 8 * 3 * 9
    assert prof.has_result()
    prof.end_profile()

    class Att(nn.Module):
        def forward(self, q):
            return torch.nn.functional.scaled_dot_product_attention(q, q, q)

    a = Att()
    _, macs, _ = get_model_profile(a, args=(torch.randn(1, 2, 8, 4), ), print_profile=False, as_string=False)
    assert macs >= 2 * 8 * 8 * 4 * 2
