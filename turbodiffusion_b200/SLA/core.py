"""Mirror of turbodiffusion/SLA/core.py: SparseLinearAttention / SageSparseLinearAttention on B200.

Same constructor arguments, parameters (`proj_l`, fp32, zero-init) and call convention (q,k,v in [B, L, H, D], result in
the same layout and dtype; SLA/core.py:181-183,253).  The forward is five kernel launches:
  quant_qk (key mean, pooled means, INT8 Q / smoothed K)  ->  block_map (pooled score, top-k, LUT)
  -> linear_moments (phi(K)^T V, sum phi(K))  ->  [tiny cuBLAS bmm: proj_l folded into the moment matrix]
  -> attn_fwd (sparse INT8 QK^T, softmax, PV, linear branch, merge).
Inputs are consumed in place in their [B, L, H, D] layout; no transposed copies of q/k/v are made.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from .._lib import DTYPE_TAG, check, lib, ptr, require_cuda, stream_ptr
from .utils import block_map_from_pools, get_block_map, get_cuda_arch, quant_qk  # noqa: F401

SAGESLA_ENABLED = True   # the INT8 path is built in (no SpargeAttn dependency)
SAGE2PP_ENABLED = True


def linear_moments(k: torch.Tensor, v: torch.Tensor):
    """kv [B,H,D(v),D(k)] fp32, ksum [B,H,D] fp32 of the local keys (accumulated: add shards / all-reduce)."""
    b, l, h, d = k.shape
    kv = torch.zeros(b, h, d, d, dtype=torch.float32, device=k.device)
    ksum = torch.zeros(b, h, d, dtype=torch.float32, device=k.device)
    check(lib().tdb200_sla_linear_moments(ptr(k), ptr(v), DTYPE_TAG[k.dtype], b, l, h, d, ptr(kv), ptr(ksum),
                                          stream_ptr(k.device)), "sla_linear_moments")
    return kv, ksum


ATTN_IMPL = os.environ.get("TDB200_ATTN_IMPL", "v2")  # "v1": one CTA per query block (round 1); "v2": persistent kernel


ATTN_TIMER = None  # bench.py installs a callable(h, mblk, topk, d) -> context manager to time the fused-attention launches


def attn_fwd(prep, v, q, lut, topk, kvw, ksum, proj_b, sm_scale, lk=None, impl=None, feature=0):
    if ATTN_TIMER is not None:
        with ATTN_TIMER(q.shape[2] * q.shape[0], (q.shape[1] + 127) // 128, topk, q.shape[3]):
            return _attn_fwd(prep, v, q, lut, topk, kvw, ksum, proj_b, sm_scale, lk, impl, feature)
    return _attn_fwd(prep, v, q, lut, topk, kvw, ksum, proj_b, sm_scale, lk, impl, feature)


def _attn_fwd(prep, v, q, lut, topk, kvw, ksum, proj_b, sm_scale, lk=None, impl=None, feature=0):
    b, l, h, d = q.shape
    lk = v.shape[1] if lk is None else lk
    out = torch.empty_like(q)
    if (impl or ATTN_IMPL) == "v1":
        assert feature == 0
        check(lib().tdb200_sla_attn_fwd(ptr(prep.q_i8), ptr(prep.q_scale), ptr(prep.k_i8), ptr(prep.k_scale), ptr(v),
                                        ptr(q), DTYPE_TAG[q.dtype], ptr(lut), topk, ptr(kvw), ptr(ksum), ptr(proj_b),
                                        ptr(out), b, l, lk, h, d, float(sm_scale), stream_ptr(q.device)), "sla_attn_fwd")
    else:
        check(lib().tdb200_sla_attn_fwd_v2(ptr(prep.q_i8), ptr(prep.q_scale), ptr(prep.k_i8), ptr(prep.k_scale), ptr(v),
                                           ptr(q), DTYPE_TAG[q.dtype], ptr(lut), topk, ptr(kvw), ptr(ksum), ptr(proj_b),
                                           ptr(out), b, l, lk, h, d, float(sm_scale), feature, stream_ptr(q.device)),
              "sla_attn_fwd")
    return out


class _SLABase(nn.Module):
    def __init__(self, head_dim, topk, feature_map="softmax", use_bf16=True, tie_feature_map_qk=True):
        super().__init__()
        self.dtype = torch.bfloat16 if use_bf16 else torch.float16
        self.topk = topk
        self.proj_l = nn.Linear(head_dim, head_dim, dtype=torch.float32)
        if feature_map != "softmax":
            # the reference also offers 'elu' / 'relu' (SLA/core.py:57-73); TurboDiffusion's released models and
            # inference scripts use the default softmax map, which is what the fused kernel implements.
            raise NotImplementedError(f"Not supported feature map {feature_map}.")
        self.feature_map_q = self.feature_map_k = lambda x: F.softmax(x, dim=-1)
        self.init_weights_()

    def init_weights_(self):
        with torch.no_grad():
            nn.init.zeros_(self.proj_l.weight)
            nn.init.zeros_(self.proj_l.bias)

    def forward(self, q, k, v, return_sparsity=False):
        """q,k,v [B, L, H, D] (any float dtype), D in {64, 128} (SLA/core.py:207) -> [B, L, H, D] in q.dtype."""
        require_cuda(q, k, v)
        dtype = q.dtype
        q, k, v = (t.to(self.dtype).contiguous() for t in (q, k, v))
        b, l, h, d = q.shape
        if d == 64:
            return self._forward_d64(q, k, v, dtype, return_sparsity)
        if d != 128:
            raise AssertionError("headdim should be in [64, 128].")  # SLA/core.py:207
        o, ratio = self._forward_d128(q, k, v, q, k, d ** -0.5, self.proj_l.weight, self.proj_l.bias)
        o = o.to(dtype)
        return (o, ratio) if return_sparsity else o

    def forward_with_lut(self, q, k, v):
        """forward() that also returns the block selection it used: (o, {"lut": int32 [B,H,Mblk,topk] ascending key-block
        ids, "sparse_map": int8 [B,H,Mblk,Nblk], "topk": int}).  Test/diagnostic entry point (sampled oracle parity at
        full shapes needs the kernel's own selection)."""
        self._keep_selection = {}
        try:
            o = self.forward(q, k, v)
            return o, self._keep_selection
        finally:
            self._keep_selection = None

    def _forward_d128(self, q, k, v, q_feat, k_feat, sm_scale, proj_w, proj_b):
        """q,k,v: tensors the sparse branch sees; q_feat,k_feat: tensors the softmax feature map sees (they differ only for
        padded 64-wide heads)."""
        prep = quant_qk(q, k)
        nblk = prep.nblk
        real_topk = min(nblk, int(self.topk * nblk))
        sparse_map, lut = block_map_from_pools(prep.q_pool, prep.k_pool, real_topk)
        if getattr(self, "_keep_selection", None) is not None:
            self._keep_selection.update(lut=lut, sparse_map=sparse_map, topk=real_topk)
        kv, ksum = linear_moments(k_feat, v)
        # proj_l folded into the moments: (phi(q) KV / den) W^T + b == phi(q) (W KV^T)^T / den + b ; kv is [dv, dk]
        kvw = torch.matmul(proj_w.float(), kv).to(self.dtype).contiguous()  # [B,H,d_out,d_k]
        o = attn_fwd(prep, v, q_feat, lut, real_topk, kvw, ksum, proj_b.float().contiguous(), sm_scale)
        return o, real_topk / nblk

    def _forward_d64(self, q, k, v, dtype, return_sparsity):
        """64-wide heads run through the 128-wide kernels: q/k/v are zero-padded (scores, pooled scores, Sage scales and
        P.V are unchanged by zero channels), while the tensors feeding softmax-over-D are padded with a large negative value
        so the padded channels get phi = 0; proj_l is embedded in the top-left 64x64 corner.  (Native 64-wide tiles: next.)"""
        pad = (0, 64)
        qz, kz, vz = (F.pad(t, pad) for t in (q, k, v))
        qf, kf = F.pad(q, pad, value=-3.0e4), F.pad(k, pad, value=-3.0e4)
        w = torch.zeros(128, 128, dtype=torch.float32, device=q.device)
        w[:64, :64] = self.proj_l.weight.float()
        bias = torch.zeros(128, dtype=torch.float32, device=q.device)
        bias[:64] = self.proj_l.bias.float()
        o, ratio = self._forward_d128(qz.contiguous(), kz.contiguous(), vz.contiguous(), qf.contiguous(), kf.contiguous(),
                                      64 ** -0.5, w, bias)
        o = o[..., :64].contiguous().to(dtype)
        return (o, ratio) if return_sparsity else o


class SparseLinearAttention(_SLABase):
    """SLA/core.py:38-119.  On B200 both classes run the same fused kernel with BLKQ=128, BLKK=64
    (the configuration modify_model.py:50 requests); BLKQ/BLKK are accepted for signature compatibility."""

    def __init__(self, head_dim, topk, feature_map="softmax", BLKQ=64, BLKK=64, use_bf16=True, tie_feature_map_qk=True):
        super().__init__(head_dim, topk, feature_map, use_bf16, tie_feature_map_qk)
        self.BLKQ, self.BLKK = BLKQ, BLKK


class SageSparseLinearAttention(_SLABase):
    """SLA/core.py:122-257."""

    def __init__(self, head_dim, topk, feature_map="softmax", use_bf16=True, tie_feature_map_qk=True):
        super().__init__(head_dim, topk, feature_map, use_bf16, tie_feature_map_qk)
