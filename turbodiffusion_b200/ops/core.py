"""Mirror of the reference operator API `turbodiffusion.ops` (turbodiffusion/ops/core.py), backed by libtdb200.so.

Same function / class names, constructor arguments, buffers and state-dict keys as the reference, so
`inference/modify_model.py:56-81` (replace_linear_norm) and checkpoints with `int8_weight / scale / bias` keys work
unchanged.  Additional fused entry points (one HBM pass instead of several) sit below the mirrored ones.
"""
from __future__ import annotations

import os
import weakref
from typing import Optional, Tuple

import torch
import torch.nn as nn

from .._lib import DTYPE_TAG, check, lib, ptr, require_cuda, stream_ptr
from ..turbo_diffusion_ops import quant_cuda, gemm_cuda, gemm_cuda_swizzle_bias


# ------------------------------------------------------------------------------------------------ mirrored API
def int8_quant(x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """ops/core.py:12-25."""
    return quant_cuda(x, None, None)


def int8_linear(x: torch.Tensor, w_q: torch.Tensor, w_s: torch.Tensor, bias: Optional[torch.Tensor] = None,
                **kwargs) -> torch.Tensor:
    """ops/core.py:28-57 (+ optional fused bias: T(T(acc)+bias), the value Int8Linear.forward :408-412 produces)."""
    assert w_q.dtype == torch.int8, "Weight tensor must be int8."
    shape = x.shape
    x = x.reshape(-1, shape[-1])
    if not x.is_contiguous():
        x = x.contiguous()
    m, n = x.shape[0], w_q.shape[0]
    y = torch.empty(m, n, dtype=x.dtype, device=x.device)  # the reference pre-zeroes (:53); every element is written here
    x_q, x_s = int8_quant(x)
    if bias is None:
        gemm_cuda(x_q, x_s, w_q, w_s, y)
    else:
        gemm_cuda_swizzle_bias(x_q, x_s, w_q, w_s, y, bias)
    return y.reshape(*shape[:-1], n)


def int8_linear_prequant(x_q: torch.Tensor, x_s: torch.Tensor, w_q: torch.Tensor, w_s: torch.Tensor,
                         bias: Optional[torch.Tensor], out_dtype: torch.dtype) -> torch.Tensor:
    """GEMM on an already-quantised activation (lets q/k/v share one quantisation of the same input)."""
    y = torch.empty(x_q.shape[0], w_q.shape[0], dtype=out_dtype, device=x_q.device)
    if bias is None:
        gemm_cuda(x_q, x_s, w_q, w_s, y)
    else:
        gemm_cuda_swizzle_bias(x_q, x_s, w_q, w_s, y, bias)
    return y


def _flatten_rows(x: torch.Tensor):
    assert x.is_contiguous(), "Input must be contiguous"
    assert x.dim() in (2, 3), "Input tensors must be batched (3D) or not batched (2D)"
    return x.reshape(-1, x.shape[-1])


def rmsnorm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    """ops/core.py:139-191: fp32 [M,N] or [B,M,N] in, fp32 out."""
    require_cuda(x, w)
    x2 = _flatten_rows(x)
    if x2.dtype != torch.float32:
        raise RuntimeError("rmsnorm expects an fp32 input (the module casts with x.float())")
    y = torch.empty_like(x2)
    check(lib().tdb200_rms_norm_f32(ptr(x2), ptr(w.float()), ptr(y), x2.shape[0], x2.shape[1], float(eps),
                                    stream_ptr(x.device)), "rmsnorm")
    return y.reshape(x.shape)


def layernorm(x, w, b, eps, elementwise_affine=True):
    """ops/core.py:380-386."""
    if elementwise_affine:
        assert w is not None and b is not None
    else:
        assert w is None and b is None
    require_cuda(x, w, b)
    x2 = _flatten_rows(x)
    if x2.dtype != torch.float32:
        raise RuntimeError("layernorm expects an fp32 input (the module casts with x.float())")
    y = torch.empty_like(x2)
    check(lib().tdb200_layer_norm_f32(ptr(x2), ptr(None if w is None else w.float()),
                                      ptr(None if b is None else b.float()), ptr(y), x2.shape[0], x2.shape[1],
                                      float(eps), stream_ptr(x.device)), "layernorm")
    return y.reshape(x.shape)


def cdiv(a: int, b: int):
    return (a + b - 1) // b


class Int8Linear(nn.Module):
    """ops/core.py:391-432.  Buffers: int8_weight [N,K] int8, scale [ceil(N/128), ceil(K/128)] fp32, bias [N]."""

    def __init__(self, in_features, out_features, bias=True, dtype=torch.bfloat16):
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        row_blocks = cdiv(out_features, b=128)
        col_blocks = cdiv(in_features, b=128)
        self.register_buffer("int8_weight", torch.empty((out_features, in_features), dtype=torch.int8))
        self.register_buffer("scale", torch.empty((row_blocks, col_blocks), dtype=torch.float32))
        if bias:
            self.register_buffer("bias", torch.empty(out_features, dtype=dtype))
        else:
            self.bias = None

    def forward(self, x):
        # quant + tcgen05 GEMM with the bias add fused into the epilogue (same value as `out + self.bias`)
        return int8_linear(x, self.int8_weight, self.scale, self.bias)

    @classmethod
    def from_linear(cls, original_linear: nn.Linear, quantize: bool = True):
        int8_layer = cls(original_linear.in_features, original_linear.out_features,
                         bias=original_linear.bias is not None, dtype=original_linear.weight.dtype)
        if quantize:
            w_data = original_linear.weight.data.cuda()
            int8_w, scale = int8_quant(w_data.contiguous())
            int8_layer.int8_weight = int8_layer.int8_weight.to(int8_w.device)
            int8_layer.scale = int8_layer.scale.to(int8_w.device)
            int8_layer.int8_weight.copy_(int8_w)
            int8_layer.scale.copy_(scale)
            if original_linear.bias is not None:
                int8_layer.bias = int8_layer.bias.to(int8_w.device)
                int8_layer.bias.data.copy_(original_linear.bias.data.cuda())
        return int8_layer


class FastRMSNorm(nn.Module):
    """ops/core.py:434-452.  forward == rmsnorm(x.float(), w, eps).to(x.dtype), done in ONE pass for 16-bit x."""

    def __init__(self, dim: int, eps: float = 1e-5):
        super().__init__()
        self.dim = dim
        self.eps = eps
        self.register_buffer("weight", torch.ones(dim))

    def forward(self, x):
        return fast_rmsnorm(x, self.weight, self.eps)

    @classmethod
    def from_rmsnorm(cls, original_rmsnorm):
        layer = cls(dim=original_rmsnorm.dim, eps=original_rmsnorm.eps)
        if original_rmsnorm.weight.device != torch.device("meta"):
            layer.weight.data.copy_(original_rmsnorm.weight.float().data)
        return layer


class FastLayerNorm(nn.Module):
    """ops/core.py:454-491."""

    def __init__(self, dim: int, eps: float = 1e-5, elementwise_affine: bool = False, bias: bool = True):
        super().__init__()
        self.dim = dim
        self.eps = eps
        self.elementwise_affine = elementwise_affine
        if self.elementwise_affine:
            self.register_buffer("weight", torch.empty(self.dim))
            if bias:
                self.register_buffer("bias", torch.empty(self.dim))
            else:
                self.bias = None
        else:
            self.register_parameter("weight", None)
            self.register_parameter("bias", None)

    def forward(self, x):
        return fast_layernorm(x, self.weight, self.bias, self.eps)

    @classmethod
    def from_layernorm(cls, original_layernorm):
        layer = cls(dim=original_layernorm.normalized_shape[0], eps=original_layernorm.eps,
                    elementwise_affine=False if original_layernorm.weight is None else True,
                    bias=original_layernorm.bias is not None)
        if original_layernorm.weight is not None and original_layernorm.weight.device != torch.device("meta"):
            layer.weight.data.copy_(original_layernorm.weight.data)
        if original_layernorm.bias is not None and original_layernorm.bias.device != torch.device("meta"):
            layer.bias.data.copy_(original_layernorm.bias.data)
        return layer


# ------------------------------------------------------------------------------------------------ fused entry points
def _f32(t: Optional[torch.Tensor], n: int, what: str) -> Optional[torch.Tensor]:
    """Side inputs of the fused kernels (norm weights, modulation vectors, gates, angle tables) are read through raw fp32
    pointers.  Checkpoints loaded with load_state_dict(assign=True) keep the checkpoint dtype (often bf16), so coerce to
    contiguous fp32 here (a no-op for fp32 contiguous tensors) and check the element count."""
    if t is None:
        return None
    require_cuda(t)
    if t.dtype != torch.float32 or not t.is_contiguous():
        t = t.float().contiguous()
    if t.numel() != n:
        raise ValueError(f"{what}: expected {n} elements, got {tuple(t.shape)}")
    return t


def _rows16(x: torch.Tensor):
    require_cuda(x)
    if x.dtype not in DTYPE_TAG:
        return None
    x2 = x.reshape(-1, x.shape[-1])
    if not x2.is_contiguous():
        x2 = x2.contiguous()
    return x2


def fast_rmsnorm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    """FastRMSNorm.forward (ops/core.py:441-442) in one pass: T(float(x) * rstd * w)."""
    x2 = _rows16(x)
    if x2 is None:  # fp32 input: the reference entry point
        return rmsnorm(x.float().contiguous(), w, eps).to(x.dtype)
    y = torch.empty_like(x2)
    if x2.numel() == 0:          # empty batch: nothing to launch (an empty tensor has no data pointer to hand to the C ABI)
        return y.reshape(x.shape)
    w = _f32(w, x2.shape[1], "fast_rmsnorm weight")
    check(lib().tdb200_rms_norm(ptr(x2), DTYPE_TAG[x.dtype], ptr(w), ptr(y), x2.shape[0], x2.shape[1],
                                float(eps), stream_ptr(x.device)), "fast_rmsnorm")
    return y.reshape(x.shape)


def fast_layernorm(x, w, b, eps) -> torch.Tensor:
    """FastLayerNorm.forward (ops/core.py:477-478) in one pass."""
    x2 = _rows16(x)
    if x2 is None:
        return layernorm(x.float().contiguous(), w, b, eps, w is not None).to(x.dtype)
    y = torch.empty_like(x2)
    if x2.numel() == 0:
        return y.reshape(x.shape)
    w, b = _f32(w, x2.shape[1], "fast_layernorm weight"), _f32(b, x2.shape[1], "fast_layernorm bias")
    check(lib().tdb200_layer_norm(ptr(x2), DTYPE_TAG[x.dtype], ptr(w), ptr(b), ptr(y), x2.shape[0], x2.shape[1], float(eps),
                                  stream_ptr(x.device)), "fast_layernorm")
    return y.reshape(x.shape)


def layernorm_modulate(x: torch.Tensor, scale: torch.Tensor, shift: torch.Tensor, eps: float) -> torch.Tensor:
    """(norm(x).float() * (1 + scale) + shift).type_as(x)  — rcm/networks/wan2pt1.py:404, scale/shift fp32 [dim]."""
    x2 = _rows16(x)
    assert x2 is not None, "layernorm_modulate expects a bf16/fp16 input"
    scale, shift = _f32(scale, x2.shape[1], "layernorm_modulate scale"), _f32(shift, x2.shape[1], "layernorm_modulate shift")
    y = torch.empty_like(x2)
    check(lib().tdb200_layer_norm_modulate(ptr(x2), DTYPE_TAG[x.dtype], ptr(scale), ptr(shift), ptr(y), x2.shape[0],
                                           x2.shape[1], float(eps), stream_ptr(x.device)), "layernorm_modulate")
    return y.reshape(x.shape)


def layernorm_modulate_quant(x: torch.Tensor, scale: torch.Tensor, shift: torch.Tensor, eps: float):
    """int8_quant(layernorm_modulate(x)) without materialising the 16-bit intermediate.  Returns (q, s)."""
    x2 = _rows16(x)
    assert x2 is not None
    m, n = x2.shape
    scale, shift = _f32(scale, n, "layernorm_modulate_quant scale"), _f32(shift, n, "layernorm_modulate_quant shift")
    q = torch.empty((m, n), dtype=torch.int8, device=x.device)
    s = torch.empty((cdiv(m, 128), cdiv(n, 128)), dtype=torch.float32, device=x.device)
    stats = torch.empty((2 * m,), dtype=torch.float32, device=x.device)
    check(lib().tdb200_layer_norm_modulate_quant(ptr(x2), DTYPE_TAG[x.dtype], ptr(scale), ptr(shift), ptr(q), ptr(s),
                                                 ptr(stats), m, n, float(eps), stream_ptr(x.device)),
          "layernorm_modulate_quant")
    return q, s


def gate_residual(x: torch.Tensor, y: torch.Tensor, gate: torch.Tensor) -> torch.Tensor:
    """x + y * gate.type_as(x)  — rcm/networks/wan2pt1.py:405-406, gate fp32 [dim]."""
    x2, y2 = _rows16(x), _rows16(y)
    gate = _f32(gate, x2.shape[1], "gate_residual gate")
    out = torch.empty_like(x2)
    check(lib().tdb200_gate_residual(ptr(x2), ptr(y2), ptr(gate), ptr(out), DTYPE_TAG[x.dtype], x2.shape[0],
                                     x2.shape[1], stream_ptr(x.device)), "gate_residual")
    return out.reshape(x.shape)


def gate_residual_stats(x: torch.Tensor, y: torch.Tensor, gate: Optional[torch.Tensor], eps: float):
    """(x + y * gate.type_as(x)  [gate None: x + y],  row statistics of that result for the next LayerNorm)."""
    x2, y2 = _rows16(x), _rows16(y)
    gate = _f32(gate, x2.shape[1], "gate_residual_stats gate")
    out = torch.empty_like(x2)
    stats = torch.empty((2 * x2.shape[0],), dtype=torch.float32, device=x.device)
    check(lib().tdb200_gate_residual_stats(ptr(x2), ptr(y2), ptr(gate), ptr(out), ptr(stats), DTYPE_TAG[x.dtype],
                                           x2.shape[0], x2.shape[1], float(eps), stream_ptr(x.device)), "gate_residual_stats")
    return out.reshape(x.shape), stats


def layernorm_modulate_quant_from_stats(x: torch.Tensor, stats: torch.Tensor, scale: torch.Tensor, shift: torch.Tensor):
    """The tile pass of layernorm_modulate_quant with precomputed row statistics (from gate_residual_stats)."""
    x2 = _rows16(x)
    m, n = x2.shape
    scale, shift = _f32(scale, n, "layernorm_modulate_quant_from_stats scale"), _f32(shift, n, "layernorm_modulate_quant_from_stats shift")
    stats = _f32(stats, 2 * m, "layernorm_modulate_quant_from_stats stats")
    q = torch.empty((m, n), dtype=torch.int8, device=x.device)
    s = torch.empty((cdiv(m, 128), cdiv(n, 128)), dtype=torch.float32, device=x.device)
    check(lib().tdb200_layer_norm_modulate_quant_stats(ptr(x2), DTYPE_TAG[x.dtype], ptr(stats), ptr(scale), ptr(shift),
                                                       ptr(q), ptr(s), m, n, stream_ptr(x.device)),
          "layernorm_modulate_quant_from_stats")
    return q, s


def rope_interleaved(x: torch.Tensor, angles: torch.Tensor) -> torch.Tensor:
    """rope_apply (wan2pt1.py:156-178).  x [..., L, H, D] with leading batch folded into L by the caller; angles [L, D/2]."""
    require_cuda(x, angles)
    l, h, d = x.shape[-3:]
    xc = x.contiguous()
    y = torch.empty_like(xc)
    rows = xc.numel() // (h * d)
    assert rows == angles.shape[0], "one angle row per token row"
    angles = _f32(angles, rows * (d // 2), "rope_interleaved angles")
    check(lib().tdb200_rope_interleaved(ptr(xc), DTYPE_TAG[x.dtype], ptr(angles), ptr(y), rows, h, d,
                                        stream_ptr(x.device)), "rope_interleaved")
    return y


_ROPE_TABLES = {}   # id(angles) -> (weakref to the angle tensor, its _version, (cos, sin) table [L, D/2, 2] fp32)
ROPE_TABLE = os.environ.get("TDB200_ROPE_TABLE", "1") != "0"   # "0": evaluate sin/cos inside the kernel (A/B timing)


def rope_cos_sin(angles: torch.Tensor) -> torch.Tensor:
    """(cos, sin) of an angle tensor [L, D/2] as one fp32 table [L, D/2, 2], built once per tensor (and per in-place version of
    it) and reused by every rmsnorm_rope call that is handed the same angles: all heads, q and k, every layer of a step."""
    key = id(angles)
    hit = _ROPE_TABLES.get(key)
    if hit is not None and hit[0]() is angles and hit[1] == angles._version:
        return hit[2]
    if len(_ROPE_TABLES) >= 16:
        _ROPE_TABLES.clear()
    a = angles.detach().float()
    table = torch.stack((torch.cos(a), torch.sin(a)), dim=-1).contiguous()
    _ROPE_TABLES[key] = (weakref.ref(angles), angles._version, table)
    return table


def rmsnorm_rope(x: torch.Tensor, w: torch.Tensor, angles: torch.Tensor, eps: float, heads: int) -> torch.Tensor:
    """rope_apply(norm_q(x).view(L, H, D), freqs) in one pass; x [L, H*D]."""
    require_cuda(x, w, angles)
    xc = x.contiguous()
    l, hd = xc.shape
    w = _f32(w, hd, "rmsnorm_rope weight")
    y = torch.empty_like(xc)
    if ROPE_TABLE:
        if angles.numel() != l * (hd // heads // 2):
            raise ValueError("rmsnorm_rope angles: expected [L, head_dim/2]")
        check(lib().tdb200_rms_norm_rope_table(ptr(xc), DTYPE_TAG[x.dtype], ptr(w), ptr(rope_cos_sin(angles)), ptr(y), l, heads,
                                               hd // heads, float(eps), stream_ptr(x.device)), "rmsnorm_rope")
        return y
    angles = _f32(angles, l * (hd // heads // 2), "rmsnorm_rope angles")
    check(lib().tdb200_rms_norm_rope(ptr(xc), DTYPE_TAG[x.dtype], ptr(w), ptr(angles), ptr(y), l, heads, hd // heads,
                                     float(eps), stream_ptr(x.device)), "rmsnorm_rope")
    return y


def wan_rope_angles(t: int, h: int, w: int, head_dim: int, device=None) -> torch.Tensor:
    """Angle table [t*h*w, head_dim/2] fp32 of Wan's 3-D RoPE (VideoRopePosition3DEmb.generate_embeddings,
    rcm/networks/wan2pt1.py:111-137): bands d_h = d_w = 2*(head_dim//6), d_t = head_dim - 2*d_h, theta 10000."""
    dh = dw = head_dim // 6 * 2
    dt = head_dim - 2 * dh

    def freqs(dim):
        return 1.0 / (10000.0 ** (torch.arange(0, dim, 2)[: dim // 2].float() / dim))

    seq = torch.arange(max(t, h, w)).float()
    ft, fh, fw = torch.outer(seq[:t], freqs(dt)), torch.outer(seq[:h], freqs(dh)), torch.outer(seq[:w], freqs(dw))
    out = torch.cat([ft[:, None, None, :].expand(t, h, w, -1), fh[None, :, None, :].expand(t, h, w, -1),
                     fw[None, None, :, :].expand(t, h, w, -1)], dim=-1).reshape(t * h * w, head_dim // 2).float().contiguous()
    return out if device is None else out.to(device)
