"""Size-independent properties at BASELINE.json's full sizes (Wan2.1-1.3B 480p: L = 32760, 12 heads x 128), where the CPU
oracle would take minutes: quant round trip, block-map structure, attention normalisation / masking, moment totals."""
import pytest
import torch

pytestmark = pytest.mark.gpu
L, H, D = 32760, 12, 128


def test_quant_round_trip_bound_full_size(cuda):
    """|x - q*s| <= s/2 for every element of a [32760, 8960] tensor, except the +amax element(s) of a block, which
    saturate from +128 to 127 (quant.hpp:48,157-163) and are off by exactly one step s."""
    from turbodiffusion_b200.turbo_diffusion_ops import quant_cuda
    g = torch.Generator(device="cuda").manual_seed(0)
    x = (torch.randn(L, 8960, generator=g, device=cuda) * 2).bfloat16()
    q, s = quant_cuda(x)
    assert (s > 0).all() and s.shape == (256, 70)
    s_full = s.repeat_interleave(128, 0)[:L].repeat_interleave(128, 1)[:, :8960]
    err = (x.float() - q.float() * s_full).abs() / s_full            # in quantisation steps
    assert err.max().item() <= 1.0 + 1e-3, err.max().item()
    over = err > 0.5 + 1e-3
    assert (q[over] == 127).all(), "only saturated +amax elements may be off by more than half a step"
    assert over.float().mean().item() < 1e-3


def test_block_map_structure_full_size(cuda):
    from turbodiffusion_b200.SLA.utils import block_map_from_pools, quant_qk
    g = torch.Generator(device="cuda").manual_seed(1)
    q = torch.randn(1, L, H, D, generator=g, device=cuda).bfloat16()
    k = torch.randn(1, L, H, D, generator=g, device=cuda).bfloat16()
    prep = quant_qk(q, k)
    assert (prep.mblk, prep.nblk) == (256, 512)
    topk = int(0.1 * prep.nblk)
    sparse_map, lut = block_map_from_pools(prep.q_pool, prep.k_pool, topk)
    assert (sparse_map.sum(-1) == topk).all()
    assert (lut[..., 1:] > lut[..., :-1]).all() and lut.min() >= 0 and lut.max() < prep.nblk
    rebuilt = torch.zeros_like(sparse_map)
    rebuilt.scatter_(-1, lut.long(), 1)
    assert torch.equal(rebuilt, sparse_map)
    # Sage codes use the whole int8 range in every block (scale = amax/127)
    full = prep.q_i8[:, :, : 255 * 128].reshape(1, H, 255, 128 * 128).to(torch.int16).abs().amax(-1)
    assert (full == 127).all()


def test_attention_rows_are_convex_combinations_full_size(cuda):
    """With V == 1 (and proj_l == 0) every output element must be exactly 1: the softmax weights of the selected key
    blocks sum to one for all 32760 x 12 rows, including the ragged last query block (120 rows) and key block (56 rows)."""
    from turbodiffusion_b200.SLA import SageSparseLinearAttention
    g = torch.Generator(device="cuda").manual_seed(2)
    q = torch.randn(1, L, H, D, generator=g, device=cuda).bfloat16()
    k = torch.randn(1, L, H, D, generator=g, device=cuda).bfloat16()
    v = torch.ones(1, L, H, D, device=cuda, dtype=torch.bfloat16)
    mod = SageSparseLinearAttention(D, 0.1).to(cuda)   # proj_l is zero-initialised (SLA/core.py:163-166)
    out = mod(q, k, v)
    assert out.shape == (1, L, H, D)
    assert (out.float() - 1.0).abs().max().item() <= 2.0 ** -7


def test_linear_moment_totals_full_size(cuda):
    """sum_dk ksum == L (each phi(k) row sums to 1 up to bf16 rounding) and kv rows sum accordingly when V == 1."""
    from turbodiffusion_b200.SLA.core import linear_moments
    g = torch.Generator(device="cuda").manual_seed(3)
    k = torch.randn(1, L, H, D, generator=g, device=cuda).bfloat16()
    v = torch.ones(1, L, H, D, device=cuda, dtype=torch.bfloat16)
    kv, ksum = linear_moments(k, v)
    assert (ksum.sum(-1) / L - 1).abs().max().item() < 2e-3
    assert (kv - ksum[:, :, None, :]).abs().max().item() <= 1e-3 * ksum.abs().max().item() + 1e-2
