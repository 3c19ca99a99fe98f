"""tcgen05 operand-convention probe: K-major A x MN-major B (bf16) — the P.V layout of the SLA attention kernel."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_umma_bf16_kmajor_a_mnmajor_b(cuda):
    from turbodiffusion_b200._lib import check, lib, ptr, stream_ptr
    g = torch.Generator().manual_seed(0)
    a = torch.randn(128, 64, generator=g).bfloat16()
    b = torch.randn(64, 128, generator=g).bfloat16()
    ref = a.float() @ b.float()
    d = torch.zeros(128, 128, device=cuda)
    a_d, b_d = a.to(cuda), b.to(cuda)
    check(lib().tdb200_selftest_umma_bf16(ptr(a_d), ptr(b_d), ptr(d), stream_ptr(cuda)), "selftest")
    torch.cuda.synchronize()
    err = (d.cpu() - ref).abs().max().item()
    assert err < 1e-3, f"max abs err {err}"
