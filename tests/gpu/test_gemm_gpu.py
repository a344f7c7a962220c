"""GPU tier: the tcgen05 GEMM against an fp32 reference."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(128, 256, 64), (256, 512, 128), (384, 768, 320), (1000, 1000, 264),
                                   (2048, 6144, 4096), (4096, 4096, 14336)])
def test_gemm_nt_matches_fp32(shape):
    from deepspeed_b200.ops.kernels import gemm_sm100
    M, N, K = shape
    torch.manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    b = torch.randn(N, K, device="cuda", dtype=torch.bfloat16)
    c = gemm_sm100.matmul_nt(a, b)
    torch.cuda.synchronize()
    ref = a.float() @ b.float().t()
    err = (c.float() - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err < 1e-2 * scale + 1e-2, (err, scale)
    lib = torch.matmul(a, b.t())
    assert (c.float() - lib.float()).abs().max().item() < 2e-2 * scale + 1e-2


def test_gemm_strided_views_and_fewer_sms():
    from deepspeed_b200.ops.kernels import gemm_sm100
    torch.manual_seed(1)
    big_a = torch.randn(512, 1024, device="cuda", dtype=torch.bfloat16)
    big_b = torch.randn(1024, 1024, device="cuda", dtype=torch.bfloat16)
    a, b = big_a[:, 128:640], big_b[:, 256:768]  # row stride 1024, K = 512
    assert gemm_sm100.supports(a, b)
    c = gemm_sm100.matmul_nt(a, b, sms=7)
    ref = a.float() @ b.float().t()
    assert (c.float() - ref).abs().max().item() < 1e-2 * ref.abs().max().item() + 1e-2


@pytest.mark.parametrize("shape", [(256, 256, 64), (128, 256, 64), (512, 512, 128), (384, 768, 320), (1000, 1000, 264),
                                   (2048, 6144, 4096), (4096, 4096, 14336), (8192, 28672, 4096)])
def test_gemm_2cta_matches_fp32(shape):
    """CTA-pair (tcgen05.mma.cta_group::2) kernel, including M/N edges that leave one CTA of a pair without rows."""
    from deepspeed_b200.ops.kernels import gemm_sm100
    M, N, K = shape
    torch.manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    b = torch.randn(N, K, device="cuda", dtype=torch.bfloat16)
    c = gemm_sm100.matmul_nt_2cta(a, b)
    torch.cuda.synchronize()
    lib = torch.matmul(a, b.t())
    scale = lib.float().abs().max().item()
    assert (c.float() - lib.float()).abs().max().item() < 2e-2 * scale + 1e-2
    # repeated launches (barrier phases, TMEM double buffering across many tiles) stay correct
    for _ in range(3):
        c2 = gemm_sm100.matmul_nt_2cta(a, b)
    assert torch.equal(c, c2)
