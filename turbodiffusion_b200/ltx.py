"""LTX-2 (TurboT2AV) prologue operators on B200 — the second client of the operator API.

Same signatures as the fused helpers in ltx_core/model/transformer/transformer.py:46-94 and the Triton fast path
(ltx_distillation/fast_norm_kernels.py): x [B, T, N] bf16/fp16, scale_shift_table [num_ada, N], timestep [B, Tt, num_ada*N]
with Tt in {1, T}."""
from __future__ import annotations

import torch

from ._lib import DTYPE_TAG, check, lib, ptr, require_cuda, stream_ptr


def _prep(x, table, timestep, num_ada):
    require_cuda(x, table, timestep)
    assert x.dim() == 3 and x.dtype in DTYPE_TAG
    b, t, n = x.shape
    ts = timestep.reshape(b, -1, num_ada * n).to(x.dtype).contiguous()
    return x.contiguous(), table.float().contiguous(), ts, b, t, ts.shape[1], n


def modulated_rms_norm_from_ada(x, scale_shift_table, timestep, scale_index, shift_index, num_ada_params, eps):
    x, tab, ts, b, t, tt, n = _prep(x, scale_shift_table, timestep, num_ada_params)
    y = torch.empty_like(x)
    check(lib().tdb200_ltx_modulated_rms_norm_ada(ptr(x), DTYPE_TAG[x.dtype], ptr(tab), ptr(ts), scale_index, shift_index,
                                                  num_ada_params, ptr(y), b, t, tt, n, float(eps), stream_ptr(x.device)),
          "ltx_modulated_rms_norm_ada")
    return y


def modulate_from_ada(x, scale_shift_table, timestep, scale_index, shift_index, num_ada_params):
    x, tab, ts, b, t, tt, n = _prep(x, scale_shift_table, timestep, num_ada_params)
    y = torch.empty_like(x)
    check(lib().tdb200_ltx_modulate_ada(ptr(x), DTYPE_TAG[x.dtype], ptr(tab), ptr(ts), scale_index, shift_index,
                                        num_ada_params, ptr(y), b, t, tt, n, stream_ptr(x.device)), "ltx_modulate_ada")
    return y


def gated_residual_from_ada(x, residual, scale_shift_table, timestep, gate_index, num_ada_params):
    x, tab, ts, b, t, tt, n = _prep(x, scale_shift_table, timestep, num_ada_params)
    residual = residual.to(x.dtype).contiguous()
    y = torch.empty_like(x)
    check(lib().tdb200_ltx_gated_residual_ada(ptr(x), ptr(residual), DTYPE_TAG[x.dtype], ptr(tab), ptr(ts), gate_index,
                                              num_ada_params, ptr(y), b, t, tt, n, stream_ptr(x.device)),
          "ltx_gated_residual_ada")
    return y


def apply_split_rotary_emb(x, cos_freqs, sin_freqs):
    """x [B, T, H*D] (or [B, T, H, D]); cos/sin [B, H, T, D/2]  (ltx_core/model/transformer/rope.py:42-60)."""
    require_cuda(x, cos_freqs, sin_freqs)
    b, h, t, half = cos_freqs.shape
    xc = x.contiguous()
    y = torch.empty_like(xc)
    c, s = cos_freqs.to(x.dtype).contiguous(), sin_freqs.to(x.dtype).contiguous()
    check(lib().tdb200_ltx_split_rope(ptr(xc), ptr(c), ptr(s), DTYPE_TAG[x.dtype], ptr(y), b, t, h, 2 * half,
                                      stream_ptr(x.device)), "ltx_split_rope")
    return y
