"""Diffusers / CLIP wrappers and the per-family fused-layer classes (host tier: graphs disabled -> eager)."""
from types import SimpleNamespace

import torch
from torch import nn


class _FakeUNet(nn.Module):

    def __init__(self):
        super().__init__()
        self.conv = nn.Conv2d(4, 4, 3, padding=1)
        self.in_channels = 4
        self.config = SimpleNamespace(in_channels=4)
        self.device, self.dtype = torch.device("cpu"), torch.float32
        self.calls = []

    def forward(self, sample, timestep, encoder_hidden_states, return_dict=True, **kw):
        self.calls.append(sorted(kw))
        return (self.conv(sample) + encoder_hidden_states.mean() + timestep, )


class _FakeVAE(nn.Module):

    def __init__(self):
        super().__init__()
        self.enc, self.dec = nn.Conv2d(3, 4, 1), nn.Conv2d(4, 3, 1)
        self.config = SimpleNamespace()

    def encode(self, x, return_dict=True):
        return (self.enc(x), )

    def decode(self, z, return_dict=True):
        return (self.dec(z), )

    def forward(self, x, return_dict=True):
        return self.decode(self.encode(x)[0])


def test_unet_vae_wrappers_forward_and_flags():
    from deepspeed_b200.model_implementations import DSUNet, DSVAE
    u = _FakeUNet()
    w = DSUNet(u, enable_cuda_graph=True)  # no CUDA here: must fall back to eager transparently
    x, t, h = torch.randn(2, 4, 8, 8), torch.tensor(3.0), torch.randn(2, 5, 16)
    out = w(x, t, h)
    assert torch.allclose(out[0], u(x, t, h)[0]) and w.fwd_count == 1 and w.in_channels == 4
    w(x, t, h, cross_attention_kwargs={"scale": 1.0}, added_cond_kwargs={"a": torch.ones(1)})
    assert u.calls[-1] == ["added_cond_kwargs", "cross_attention_kwargs"]
    assert not any(p.requires_grad for p in u.parameters())
    v = DSVAE(_FakeVAE(), enable_cuda_graph=False)
    img = torch.randn(1, 3, 4, 4)
    z = v.encode(img)[0]
    assert z.shape == (1, 4, 4, 4) and v.decode(z)[0].shape == img.shape and v(img)[0].shape == img.shape


def test_graphed_callable_signature_cache_is_eager_on_cpu():
    from deepspeed_b200.model_implementations.features.cuda_graph import GraphedCallable, _sig
    g = GraphedCallable(lambda a, scale=1.0: a * scale, enabled=True)
    assert g.enabled == torch.cuda.is_available()
    assert torch.equal(g(torch.ones(3), scale=2.0), torch.full((3, ), 2.0))
    assert _sig((torch.ones(2, 3), 1)) != _sig((torch.ones(3, 2), 1)) and _sig({"a": 1}) == _sig({"a": 1})


def test_clip_wrapper_installs_mask_builder():
    from deepspeed_b200.model_implementations import DSClipEncoder

    class TM(nn.Module):

        def _build_causal_attention_mask(self, b, s, dt):
            raise AssertionError("should have been replaced")

    class Enc(nn.Module):

        def __init__(self):
            super().__init__()
            self.text_model = TM()
            self.device, self.dtype, self.config = torch.device("cpu"), torch.float32, SimpleNamespace()

        def forward(self, ids):
            m = self.text_model._build_causal_attention_mask(ids.shape[0], ids.shape[1], torch.float32)
            return m

    e = DSClipEncoder(Enc())
    m = e(torch.zeros(2, 5, dtype=torch.long))
    assert m.shape == (2, 1, 5, 5) and m[0, 0, 0, 1] == torch.finfo(torch.float32).min and m[0, 0, 1, 0] == 0


def test_family_layer_classes_share_the_fused_layer():
    from deepspeed_b200 import model_implementations as MI
    from deepspeed_b200.ops.transformer.inference.config import DeepSpeedInferenceConfig
    for cls in (MI.DeepSpeedBERTInference, MI.DeepSpeedBloomInference, MI.DeepSpeedGPTInference, MI.DeepSpeedLlama2Inference,
                MI.DeepSpeedMegatronGPTInference, MI.DeepSpeedOPTInference):
        assert issubclass(cls, MI.DeepSpeedTransformerInference)
    cfg = DeepSpeedInferenceConfig(hidden_size=32, intermediate_size=64, heads=4, num_hidden_layers=1, dtype=torch.float32,
                                   pre_layer_norm=True, max_out_tokens=16)
    layer = MI.DeepSpeedGPTInference(cfg)
    for p in layer.parameters():
        torch.nn.init.normal_(p, std=0.05)
    y = layer(torch.randn(1, 4, 32))
    y = y[0] if isinstance(y, tuple) else y
    assert y.shape == (1, 4, 32) and torch.isfinite(y).all()
