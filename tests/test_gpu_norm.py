"""GPU parity: FastNorm family, AdaLN modulation / gate, RoPE against the CPU oracle (<= 1 ulp of the output type)."""
import pytest
import torch

from oracle import td_oracle as O

pytestmark = pytest.mark.gpu


def _ulp_report(got, ref, floor_frac=2e-2):
    """Fraction of elements that differ and the worst difference in ulps of the 16-bit type.  The ulp is taken at
    max(|ref|, floor_frac * max|ref|): results that are a small difference of O(1) terms (x - mean, h*(1+s)+t) carry
    the absolute rounding error of those terms, so a purely relative ulp near zero would be meaningless."""
    g, r = got.float(), ref.float()
    neq = g != r
    if not neq.any():
        return 0.0, 0.0
    mag = torch.maximum(r.abs(), floor_frac * r.abs().max())
    ulp = mag.log2().floor().exp2() * (2.0 ** -7 if ref.dtype == torch.bfloat16 else 2.0 ** -10)
    return neq.float().mean().item(), ((g - r).abs() / ulp).max().item()


def _x(m, n, seed, dtype=torch.bfloat16, shift=0.3):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(m, n, generator=g) * 1.7 + shift).to(dtype)


DIMS = [(64, 1536), (17, 5120), (33, 128), (5, 4096), (3, 8960), (2, 13824)]


@pytest.mark.parametrize("m,n", DIMS)
def test_fast_rmsnorm(cuda, m, n):
    import turbodiffusion_b200.ops as ops
    x = _x(m, n, n + m)
    w = torch.rand(n) + 0.5
    ref = O.fast_rmsnorm(x, w, 1e-6)
    mod = ops.FastRMSNorm(n, eps=1e-6).to(cuda)
    mod.weight.copy_(w)
    got = mod(x.to(cuda)).cpu()
    frac, worst = _ulp_report(got, ref)
    assert worst <= 1.0 and frac < 2e-3, (frac, worst)


@pytest.mark.parametrize("m,n", DIMS)
@pytest.mark.parametrize("affine", [False, True])
def test_fast_layernorm_reproduces_reference_variance_padding(cuda, m, n, affine):
    import turbodiffusion_b200.ops as ops
    x = _x(m, n, 7 * n + m, shift=1.5)  # non-zero mean makes the (N2-N)*mean^2 term visible
    w = torch.rand(n) + 0.5 if affine else None
    b = torch.randn(n) if affine else None
    ref = O.fast_layernorm(x, w, b, 1e-6)
    mod = ops.FastLayerNorm(n, eps=1e-6, elementwise_affine=affine).to(cuda)
    if affine:
        mod.weight.copy_(w)
        mod.bias.copy_(b)
    got = mod(x.to(cuda)).cpu()
    frac, worst = _ulp_report(got, ref)
    assert worst <= 1.0 and frac < 2e-3, (frac, worst)
    if O.next_pow2(n) != n:
        textbook = O.layernorm_f32(x, w, b, 1e-6, reference_padding_quirk=False).to(x.dtype)
        assert (got.float() - textbook.float()).abs().max() > (got.float() - ref.float()).abs().max()


def test_fp32_reference_entry_points(cuda):
    from turbodiffusion_b200 import turbo_diffusion_ops as tdo
    import turbodiffusion_b200.ops as ops
    x = _x(37, 1536, 1, torch.float32)
    w, b = torch.rand(1536) + 0.5, torch.randn(1536)
    for got, ref in (
        (tdo.rms_norm_cuda(x.to(cuda), 1e-6, w.to(cuda)), O.rmsnorm_f32(x, w, 1e-6)),
        (tdo.layer_norm_cuda(x.to(cuda), 1e-6, w.to(cuda), b.to(cuda)), O.layernorm_f32(x, w, b, 1e-6)),
        (ops.rmsnorm(x.to(cuda).reshape(1, 37, 1536), w.to(cuda), 1e-6).reshape(37, 1536), O.rmsnorm_f32(x, w, 1e-6)),
        (ops.layernorm(x.to(cuda), None, None, 1e-6, False), O.layernorm_f32(x, None, None, 1e-6)),
    ):
        torch.testing.assert_close(got.cpu(), ref, rtol=2e-5, atol=2e-5)


@pytest.mark.parametrize("m,n", [(300, 1536), (131, 5120), (128, 256)])
def test_layernorm_modulate_and_fused_quant(cuda, m, n):
    import turbodiffusion_b200.ops as ops
    x = _x(m, n, 3 * n + m, shift=0.7)
    g = torch.Generator().manual_seed(9)
    scale, shift = torch.randn(n, generator=g) * 0.1, torch.randn(n, generator=g) * 0.1
    ref = O.ln_modulate(x, scale, shift, 1e-6)
    xd, sd, hd = x.to(cuda), scale.to(cuda), shift.to(cuda)
    got = ops.layernorm_modulate(xd, sd, hd, 1e-6)
    # The LayerNorm value is rounded to bf16 BEFORE the modulation (wan2pt1.py:404).  A 1-ulp flip of that intermediate
    # (summation order of the row statistics) moves the result by ulp(h)*|1+scale|, which can be many ulps of a result
    # that cancels against the shift.  Bound = one ulp of the intermediate propagated + one ulp of the result.
    h = O.fast_layernorm(x, None, None, 1e-6).float()
    ulp = lambda t: torch.maximum(t.abs(), torch.tensor(1e-30)).log2().floor().exp2() * 2.0 ** -7
    bound = ulp(h) * (1 + scale).abs() + ulp(ref.float())
    err = (got.cpu().float() - ref.float()).abs()
    assert (err <= bound).all(), (err / bound).max()
    assert (err > 0).float().mean() < 2e-3
    # the fused quant must equal quantising the kernel's own 16-bit output: bit-exact
    q_ref, s_ref = O.int8_quant(got.cpu())
    q, s = ops.layernorm_modulate_quant(xd, sd, hd, 1e-6)
    assert torch.equal(s.cpu(), s_ref) and torch.equal(q.cpu(), q_ref)


def test_gate_residual_bit_exact(cuda):
    import turbodiffusion_b200.ops as ops
    x, y = _x(257, 1536, 1), _x(257, 1536, 2)
    gate = torch.randn(1536)
    ref = O.gate_residual(x, y, gate)
    got = ops.gate_residual(x.to(cuda), y.to(cuda), gate.to(cuda)).cpu()
    assert torch.equal(got.view(torch.int16), ref.view(torch.int16))


@pytest.mark.parametrize("h,d", [(12, 128), (24, 64)])
def test_rope_and_fused_rmsnorm_rope(cuda, h, d):
    import turbodiffusion_b200.ops as ops
    t, hh, ww = 3, 6, 10
    l = t * hh * ww
    ang = O.wan_rope_angles(t, hh, ww, d)
    x = _x(l, h * d, 4).reshape(l, h, d)
    ref = O.rope_interleaved(x, ang)
    got = ops.rope_interleaved(x.to(cuda), ang.to(cuda)).cpu()
    frac, worst = _ulp_report(got, ref)
    assert worst <= 1.0 and frac < 5e-3, (frac, worst)
    w = torch.rand(h * d, generator=torch.Generator().manual_seed(17)) + 0.5
    ref2 = O.rms_norm_rope(x, w, ang, 1e-6)
    got2 = ops.rmsnorm_rope(x.reshape(l, h * d).to(cuda), w.to(cuda), ang.to(cuda), 1e-6, h).cpu().reshape(l, h, d)
    frac, worst = _ulp_report(got2, ref2)
    # the rotation reads the T-ROUNDED norm output: where the fp32 norm value sits on a rounding boundary the kernel's and the
    # oracle's summation orders round it to neighbouring T values, which moves the rotated pair by up to |cos|+|sin| <= 1.42
    # ulps of the INPUT magnitude -- more than one ulp of a smaller (cancelled) output.  Seen on ~1e-4 of the elements.
    assert worst <= 2.0 and frac < 5e-3, (frac, worst)


@pytest.mark.parametrize("m,n,gated", [(300, 1536, True), (131, 5120, False)])
def test_residual_with_fused_stats_equals_the_two_pass_path(cuda, m, n, gated):
    """gate_residual_stats + layernorm_modulate_quant_from_stats == gate_residual + layernorm_modulate_quant, bit for bit."""
    import turbodiffusion_b200.ops as ops
    x, y = _x(m, n, 1).to(cuda), _x(m, n, 2).to(cuda)
    g = torch.Generator().manual_seed(3)
    gate = torch.randn(n, generator=g).to(cuda) if gated else None
    scale, shift = (torch.randn(n, generator=g) * 0.1).to(cuda), (torch.randn(n, generator=g) * 0.1).to(cuda)
    out_a = ops.gate_residual(x, y, gate) if gated else x + y
    q_a, s_a = ops.layernorm_modulate_quant(out_a, scale, shift, 1e-6)
    out_b, stats = ops.gate_residual_stats(x, y, gate, 1e-6)
    q_b, s_b = ops.layernorm_modulate_quant_from_stats(out_b, stats, scale, shift)
    torch.cuda.synchronize()
    assert torch.equal(out_a.view(torch.int16), out_b.view(torch.int16))
    assert torch.equal(s_a, s_b) and torch.equal(q_a, q_b)


def test_rmsnorm_rope_table_and_in_kernel_sincos_agree_and_the_table_follows_the_angles(cuda, monkeypatch):
    """ops.rmsnorm_rope reads (cos, sin) from a table cached per angle tensor (tdb200_rms_norm_rope_table); TDB200_ROPE_TABLE=0
    evaluates sin/cos in the kernel.  Both sit within the fused-RoPE tolerance of the oracle; an in-place change of the angle
    tensor or a new tensor must not be served a stale table."""
    import turbodiffusion_b200.ops as ops
    from turbodiffusion_b200.ops import core as C
    h, d, (t, hh, ww) = 12, 128, (3, 6, 10)
    l = t * hh * ww
    ang = O.wan_rope_angles(t, hh, ww, d)
    x = _x(l, h * d, 4)
    w = torch.rand(h * d, generator=torch.Generator().manual_seed(17)) + 0.5
    ref = O.rms_norm_rope(x.reshape(l, h, d), w, ang, 1e-6)
    ang_dev = ang.to(cuda)
    outs = {}
    for table in (True, False):
        monkeypatch.setattr(C, "ROPE_TABLE", table)
        outs[table] = ops.rmsnorm_rope(x.to(cuda), w.to(cuda), ang_dev, 1e-6, h).cpu().reshape(l, h, d)
        frac, worst = _ulp_report(outs[table], ref)
        assert worst <= 2.0 and frac < 5e-3, (table, frac, worst)
    assert (outs[True] != outs[False]).float().mean().item() < 5e-3
    monkeypatch.setattr(C, "ROPE_TABLE", True)
    ang_dev.mul_(0.5)                                     # in place: the cached table is stale now
    got = ops.rmsnorm_rope(x.to(cuda), w.to(cuda), ang_dev, 1e-6, h).cpu().reshape(l, h, d)
    frac, worst = _ulp_report(got, O.rms_norm_rope(x.reshape(l, h, d), w, ang * 0.5, 1e-6))
    assert worst <= 2.0 and frac < 5e-3, (frac, worst)
    got = ops.rmsnorm_rope(x.to(cuda), w.to(cuda), torch.zeros_like(ang_dev), 1e-6, h).cpu()     # zero angles: identity rotation
    assert torch.equal(got, ops.fast_rmsnorm(x.to(cuda), w.to(cuda), 1e-6).cpu())
