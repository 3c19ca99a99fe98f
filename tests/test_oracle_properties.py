"""Size-independent properties of the CPU oracle and of the host helpers (hypothesis-driven where the input space is
large).  These guard the checker itself: the GPU parity tests are only as good as the oracle they compare against."""
import os
import sys

import pytest
import torch
from hypothesis import given, settings, strategies as st

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import td_oracle as O  # noqa: E402
from turbodiffusion_b200.dist import shard_rows, split_heads  # noqa: E402


@settings(max_examples=25, deadline=None)
@given(st.integers(1, 300), st.integers(1, 300), st.integers(0, 2 ** 31 - 1), st.sampled_from([torch.bfloat16, torch.float16]))
def test_int8_quant_round_trip_bound(m, k, seed, dtype):
    """ops/quant/quant.hpp:86-99,122-164: s = amax/128 per 128x128 block; |x - q*s| <= s/2 except the +amax element(s),
    which saturate from +128 to 127 (error exactly s); ragged edges never read outside the tensor."""
    g = torch.Generator().manual_seed(seed)
    x = (torch.randn(m, k, generator=g) * 3).to(dtype)
    q, s = O.int8_quant(x)
    assert q.shape == (m, k) and s.shape == (O.cdiv(m, 128), O.cdiv(k, 128)) and (s > 0).all()
    s_full = s.repeat_interleave(128, 0)[:m].repeat_interleave(128, 1)[:, :k]
    err = (x.float() - q.float() * s_full).abs() / s_full
    assert err.max() <= 1.0 + 1e-3
    over = err > 0.5 + 1e-3
    assert (q[over] == 127).all()
    # scale is exactly amax/128 of its block
    blk = x.float()[:128, :128].abs().max().clamp_min(1e-8) / 128.0
    assert s[0, 0].item() == blk.item()


@settings(max_examples=15, deadline=None)
@given(st.integers(1, 200), st.integers(1, 40), st.integers(1, 3), st.integers(0, 2 ** 31 - 1))
def test_w8a8_gemm_is_the_fp32_fma_chain(m, n8, kb, seed):
    """kernel.hpp:391-427 + utils.hpp:116-121: per 128-wide K block acc = fma(float(int32 partial), sA*sB, acc), ascending;
    the C and the torch restatements agree bit for bit, and with one K block it is a single rounded product."""
    g = torch.Generator().manual_seed(seed)
    n, k = n8 * 8, kb * 128
    a = torch.randint(-128, 128, (m, k), generator=g, dtype=torch.int8)
    b = torch.randint(-128, 128, (n, k), generator=g, dtype=torch.int8)
    a_s = torch.rand(O.cdiv(m, 128), kb, generator=g) * 0.02 + 1e-3
    b_s = torch.rand(O.cdiv(n, 128), kb, generator=g) * 0.02 + 1e-3
    y_c = O.int8_gemm_f32(a, a_s, b, b_s)
    y_t = O._int8_gemm_f32_torch(a, a_s, b, b_s)
    assert torch.equal(y_c, y_t)
    if kb == 1:
        part = (a.double() @ b.double().t()).float()
        sc = a_s.repeat_interleave(128, 0)[:m, :1] * b_s.repeat_interleave(128, 0)[:n, 0][None, :]
        assert torch.equal(y_c, part * sc)


@settings(max_examples=20, deadline=None)
@given(st.integers(1, 6), st.integers(2, 40), st.integers(1, 40), st.integers(0, 2 ** 31 - 1))
def test_select_topk_is_a_stable_top_set(rows, nblk, topk, seed):
    """SLA/utils.py:59-66 with the tie rule the kernel implements (lowest index wins): every row selects exactly topk
    blocks, the LUT is ascending, every selected score >= every rejected one, and among equal scores at the threshold the
    lower indices are the selected ones."""
    topk = min(topk, nblk)
    g = torch.Generator().manual_seed(seed)
    scores = torch.randint(-3, 4, (1, 1, rows, nblk), generator=g).float().bfloat16()   # few distinct values: many ties
    sparse_map, lut = O.select_topk(scores, topk)
    assert (sparse_map.sum(-1) == topk).all()
    assert (lut[..., 1:] > lut[..., :-1]).all()
    for r in range(rows):
        row, sel = scores[0, 0, r].float(), sparse_map[0, 0, r].bool()
        assert torch.equal(torch.nonzero(sel).flatten().int(), lut[0, 0, r])
        if sel.all():
            continue
        thr = row[sel].min()
        assert row[~sel].max() <= thr
        tie = row == thr
        sel_tie, rej_tie = torch.nonzero(tie & sel).flatten(), torch.nonzero(tie & ~sel).flatten()
        if len(sel_tie) and len(rej_tie):
            assert sel_tie.max() < rej_tie.min()


def test_rope_is_a_rotation_and_adds_angles():
    """wan2pt1.py:156-178: each (2i, 2i+1) pair is rotated by its angle: norms are preserved and rotating by a then b
    equals rotating by a+b (checked in fp32 through float inputs)."""
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 7, 3, 16, generator=g)
    a, b = torch.rand(7, 8, generator=g) * 6, torch.rand(7, 8, generator=g) * 6
    y = O.rope_interleaved(x, a)
    pair = lambda t: (t[..., 0::2] ** 2 + t[..., 1::2] ** 2)
    assert torch.allclose(pair(y), pair(x), rtol=1e-5, atol=1e-6)
    assert torch.allclose(O.rope_interleaved(y, b), O.rope_interleaved(x, a + b), rtol=1e-4, atol=1e-5)
    assert torch.equal(O.rope_interleaved(x, torch.zeros(7, 8)), x)


@settings(max_examples=20, deadline=None)
@given(st.integers(1, 500), st.sampled_from([64, 128]), st.integers(0, 2 ** 31 - 1))
def test_sage_quant_round_trip(l, blk, seed):
    """Sage per-block INT8 (emulation of SpargeAttn's get_vanilla_qk_quant): scale = amax/127 + 1e-7 per `blk` rows, round
    half away from zero: |x - q*scale| <= scale/2 everywhere, codes within [-127, 127]."""
    g = torch.Generator().manual_seed(seed)
    x = (torch.randn(1, 2, l, 16, generator=g) * 2).bfloat16()
    q, s = O.sage_quant_blocks(x, blk)
    assert q.shape == x.shape and s.shape == (1, 2, O.cdiv(l, blk))
    s_full = s.repeat_interleave(blk, -1)[..., :l, None]
    assert ((x.float() - q.float() * s_full).abs() <= s_full * 0.5 * (1 + 1e-5) + 1e-7).all()
    assert q.abs().max() <= 127


def test_layernorm_quirk_reduces_to_textbook_for_power_of_two_rows():
    """ops/core.py:217-224,315-322: the variance padding term (N2-N)*mean^2/N vanishes when N is a power of two."""
    g = torch.Generator().manual_seed(5)
    x = torch.randn(4, 256, generator=g) + 0.7
    ref = torch.nn.functional.layer_norm(x, (256,), eps=1e-6)
    assert torch.allclose(O.layernorm_f32(x, None, None, 1e-6), ref, rtol=1e-5, atol=1e-5)
    x = torch.randn(4, 384, generator=g) + 0.7                    # N2 = 512: the quirk term is (128/384)*mean^2
    mean = x.mean(-1, keepdim=True)
    var = ((x - mean) ** 2).mean(-1, keepdim=True) + (512 - 384) / 384 * mean ** 2
    assert torch.allclose(O.layernorm_f32(x, None, None, 1e-6), (x - mean) * torch.rsqrt(var + 1e-6), rtol=1e-5, atol=1e-5)
    textbook = torch.nn.functional.layer_norm(x, (384,), eps=1e-6)
    assert torch.allclose(O.layernorm_f32(x, None, None, 1e-6, reference_padding_quirk=False), textbook, rtol=1e-5, atol=1e-5)
    assert not torch.allclose(O.layernorm_f32(x, None, None, 1e-6), textbook, rtol=1e-3, atol=1e-3)


@settings(max_examples=60, deadline=None)
@given(st.integers(1, 200000), st.integers(1, 8))
def test_shard_rows_partition(total, world):
    """dist.shard_rows: contiguous 128-aligned shards that cover [0, total) exactly once; only trailing ranks may be short
    or empty; every rank pads to the same length."""
    covered, pads = 0, set()
    for r in range(world):
        b, e, pad = shard_rows(total, world, r)
        assert b == covered and b <= e <= total and (b % 128 == 0 or b == total) and pad % 128 == 0 and e - b <= pad
        covered = e
        pads.add(pad)
    assert covered == total and len(pads) == 1


@settings(max_examples=60, deadline=None)
@given(st.integers(1, 64), st.integers(1, 8))
def test_split_heads_partition(heads, world):
    counts, offs = split_heads(heads, world)
    assert sum(counts) == heads and offs[0] == 0 and max(counts) - min(counts) <= 1
    assert all(offs[i + 1] == offs[i] + counts[i] for i in range(world - 1)) and sorted(counts, reverse=True) == counts


def test_proj_l_folding_identity_and_shard_additivity():
    """Two algebraic facts the fused kernels rely on (SLA/core.py:243-253):
    (1) proj_l can be folded into the moments:  (phi(q) KV / den) W^T + b == phi(q) (W KV^T)^T / den + b;
    (2) the moments KV = phi(K)^T V and ksum = sum phi(K) are sums over key rows, so row shards (ranks) simply add."""
    g = torch.Generator().manual_seed(9)
    b, h, l, d = 1, 2, 300, 16
    q, k, v = (torch.randn(b, h, l, d, generator=g).double() for _ in range(3))
    w, bias = torch.randn(d, d, generator=g).double() * 0.1, torch.randn(d, generator=g).double() * 0.1
    pq, pk = torch.softmax(q, -1), torch.softmax(k, -1)
    kv = pk.transpose(-1, -2) @ v                       # [b,h,dk,dv]
    ks = pk.sum(-2, keepdim=True)
    den = 1e-5 + (pq * ks).sum(-1, keepdim=True)
    ref = ((pq @ kv) / den) @ w.t() + bias
    kvw = w @ kv.transpose(-1, -2)                      # [b,h,d_out,dk]: the K-major B operand the kernel consumes
    folded = (pq @ kvw.transpose(-1, -2)) / den + bias
    assert torch.allclose(ref, folded, rtol=1e-12, atol=1e-12)
    assert torch.allclose(O.linear_branch(q.float(), k.float(), v.float(), w.float(), bias.float(), torch.float32, exact=True).double(),
                          ref, rtol=1e-4, atol=1e-5)
    cut = 128                                           # shard boundary (multiple of the 64-row key block)
    kv2 = pk[..., :cut, :].transpose(-1, -2) @ v[..., :cut, :] + pk[..., cut:, :].transpose(-1, -2) @ v[..., cut:, :]
    assert torch.allclose(kv, kv2, rtol=1e-12, atol=1e-12)
    assert torch.allclose(ks, pk[..., :cut, :].sum(-2, keepdim=True) + pk[..., cut:, :].sum(-2, keepdim=True), rtol=1e-12)


def test_padded_64_wide_heads_equal_native_64_wide_math():
    """The d=64 path runs through the 128-wide kernels with zero-padded q/k/v (scores, pooled scores, P.V unchanged) and a
    large negative pad for the softmax feature map (phi = 0 on padded channels); in exact arithmetic both formulations of
    the module agree (turbodiffusion_b200/SLA/core.py _forward_d64)."""
    g = torch.Generator().manual_seed(4)
    bsz, l, h, d = 1, 200, 2, 64
    q, k, v = (torch.randn(bsz, l, h, d, generator=g).bfloat16() for _ in range(3))
    w, bias = torch.randn(d, d, generator=g) * 0.05, torch.randn(d, generator=g) * 0.05
    ref = O.sla_forward(q, k, v, w, bias, 0.5, mode="exact").float()
    pad = lambda t, val=0.0: torch.nn.functional.pad(t, (0, 64), value=val)
    qh, kh, vh = (pad(t).transpose(1, 2).contiguous() for t in (q, k, v))
    _, lut, _ = O.get_block_map(q.transpose(1, 2).contiguous(), k.transpose(1, 2).contiguous(), 0.5)
    _, lut_p, _ = O.get_block_map(qh, kh, 0.5)
    assert torch.equal(lut, lut_p)                                             # zero channels do not move pooled scores
    o_s = O.sparse_attention(qh, kh, vh, lut, 128, 64, sm_scale=64 ** -0.5)[..., :64]
    qf, kf = (pad(t, -3.0e4).transpose(1, 2).contiguous() for t in (q, k))
    w128 = torch.zeros(128, 128)
    w128[:64, :64] = w
    b128 = torch.zeros(128)
    b128[:64] = bias
    o_l = O.linear_branch(qf, kf, vh, w128, b128, torch.bfloat16, exact=True)[..., :64]
    got = (o_s + o_l).to(torch.bfloat16).transpose(1, 2).float()
    assert O.stats(got, ref)["rel_l2"] < 1e-6
