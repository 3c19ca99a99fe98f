"""GPU parity of the LTX-2 prologue variants (row a12) against the formulas the reference's own known-answer tests pin
(TurboT2AV/LTX-2/packages/ltx-core/tests/test_transformer_fusion_helpers.py:13-87), evaluated in fp32 on the CPU."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _inputs(b, t, n, tt, seed=7, dtype=torch.bfloat16):
    torch.manual_seed(seed)
    x = torch.randn(b, t, n).to(dtype)
    table = torch.randn(6, n)
    timestep = (torch.randn(b, tt, 6 * n) * 0.5).to(dtype)
    return x, table, timestep


def _ada(table, timestep, b, n):
    return table[None, None] + timestep.float().reshape(b, -1, 6, n)


def _close(got, ref):
    g, r = got.float().cpu(), ref
    ulp = torch.maximum(r.abs(), 2e-2 * r.abs().max()).log2().floor().exp2() * 2.0 ** -7
    assert ((g - r).abs() <= 1.01 * ulp).all(), ((g - r).abs() / ulp).max()


@pytest.mark.parametrize("b,t,n,tt", [(2, 3, 4096, 3), (1, 37, 4096, 1), (2, 5, 2048, 5)])
def test_ltx_ada_kernels(cuda, b, t, n, tt):
    from turbodiffusion_b200 import ltx
    x, table, ts = _inputs(b, t, n, tt)
    ada = _ada(table, ts, b, n)
    xf = x.float()
    xd, td, tsd = x.to(cuda), table.to(cuda), ts.to(cuda)
    rms = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-6)
    _close(ltx.modulated_rms_norm_from_ada(xd, td, tsd, 1, 0, 6, 1e-6), rms * (1 + ada[:, :, 1]) + ada[:, :, 0])
    _close(ltx.modulate_from_ada(xd, td, tsd, 2, 3, 6), xf * (1 + ada[:, :, 2]) + ada[:, :, 3])
    res = torch.randn(b, t, n).bfloat16()
    _close(ltx.gated_residual_from_ada(xd, res.to(cuda), td, tsd, 5, 6), xf + res.float() * ada[:, :, 5])


def test_ltx_split_rope(cuda):
    from turbodiffusion_b200 import ltx
    b, t, h, d = 2, 9, 4, 128
    torch.manual_seed(3)
    x = torch.randn(b, t, h * d).bfloat16()
    ang = torch.rand(b, h, t, d // 2) * 6.28
    cos, sin = ang.cos().bfloat16(), ang.sin().bfloat16()
    xs = x.float().reshape(b, t, h, d).transpose(1, 2)          # [b,h,t,d]
    x1, x2 = xs[..., : d // 2], xs[..., d // 2:]
    ref = torch.cat([x1 * cos.float() - x2 * sin.float(), x2 * cos.float() + x1 * sin.float()], -1)
    ref = ref.transpose(1, 2).reshape(b, t, h * d)
    got = ltx.apply_split_rotary_emb(x.to(cuda), cos.to(cuda), sin.to(cuda))
    _close(got, ref)


@pytest.mark.parametrize("m,n,k", [(256, 512, 256), (300, 384, 1024), (1000, 4096, 4096)])
def test_ltx_post_scale_w8a8_bit_exact(cuda, m, n, k):
    """Per-row quant + int32-accumulate GEMM with one post-scale epilogue (tilelang_w8a8.py) against the oracle."""
    from oracle import td_oracle as O
    from turbodiffusion_b200 import ltx
    torch.manual_seed(m + n + k)
    x = (torch.randn(m, k) * 2).bfloat16()
    x[:, ::61] *= 12
    w = (torch.randn(n, k) * 0.05).bfloat16()
    bias = torch.randn(n).bfloat16()
    xq_ref, xs_ref = O.ltx_row_quant_int8(x)
    wq_ref, ws_ref = O.ltx_row_quant_int8(w)
    xq, xs = ltx.row_quant_int8(x.to(cuda))
    assert torch.equal(xs.cpu(), xs_ref) and torch.equal(xq.cpu(), xq_ref)
    y_ref = O.ltx_gemm_post_scale(xq_ref, xs_ref, wq_ref, ws_ref, bias)
    y = ltx.gemm_int8_post_scale_bias(xq, xs, wq_ref.to(cuda), ws_ref.to(cuda), bias.to(cuda)).cpu()
    assert torch.equal(y.view(torch.int16), y_ref.view(torch.int16)), (y.float() - y_ref.float()).abs().max()
    lin = torch.nn.Linear(k, n).to(cuda).bfloat16()
    mod = ltx.PostScaleInt8Linear.from_linear(lin)
    out = mod(x.to(cuda).reshape(2, m // 2, k))
    ref = torch.nn.functional.linear(x.to(cuda).float(), lin.weight.float(), lin.bias.float())
    # per-row scales with 12x outlier columns in x: a few % of relative error is inherent to this quantisation scheme
    assert O.stats(out.reshape(m, n).float().cpu(), ref.cpu())["rel_l2"] < 0.15
