"""Mirror of turbodiffusion/SLA/core.py: SparseLinearAttention / SageSparseLinearAttention on B200.

Same constructor arguments, parameters (`proj_l`, fp32, zero-init) and call convention (q,k,v in [B, L, H, D], result in
the same layout and dtype; SLA/core.py:181-183,253).  The forward is five kernel launches:
  quant_qk (key mean, pooled means, INT8 Q / smoothed K)  ->  block_map (pooled score, top-k, LUT)
  -> linear_moments (phi(K)^T V, sum phi(K))  ->  [tiny cuBLAS bmm: proj_l folded into the moment matrix]
  -> attn_fwd (sparse INT8 QK^T, softmax, PV, linear branch, merge).
Inputs are consumed in place in their [B, L, H, D] layout; no transposed copies of q/k/v are made.
"""
from __future__ import annotations

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


def attn_fwd(prep, v, q, lut, topk, kvw, ksum, proj_b, sm_scale, lk=None):
    b, l, h, d = q.shape
    lk = v.shape[1] if lk is None else lk
    out = torch.empty_like(q)
    check(lib().tdb200_sla_attn_fwd(ptr(prep.q_i8), ptr(prep.q_scale), ptr(prep.k_i8), ptr(prep.k_scale), ptr(v),
                                    ptr(q), DTYPE_TAG[q.dtype], ptr(lut), topk, ptr(kvw), ptr(ksum), ptr(proj_b),
                                    ptr(out), b, l, lk, h, d, float(sm_scale), stream_ptr(q.device)), "sla_attn_fwd")
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
        """q,k,v [B, L, H, D] (any float dtype) -> [B, L, H, D] in q.dtype."""
        require_cuda(q, k, v)
        dtype = q.dtype
        q, k, v = (t.to(self.dtype).contiguous() for t in (q, k, v))
        b, l, h, d = q.shape
        prep = quant_qk(q, k)
        nblk = prep.nblk
        real_topk = min(nblk, int(self.topk * nblk))
        sparse_map, lut = block_map_from_pools(prep.q_pool, prep.k_pool, real_topk)
        kv, ksum = linear_moments(k, v)
        # proj_l folded into the moments: (phi(q) KV / den) W^T + b == phi(q) (W KV^T)^T / den + b ; kv is [dv, dk]
        kvw = torch.matmul(self.proj_l.weight.float(), kv).to(self.dtype).contiguous()  # [B,H,d_out,d_k]
        o = attn_fwd(prep, v, q, lut, real_topk, kvw, ksum, self.proj_l.bias.float().contiguous(), d ** -0.5)
        o = o.to(dtype)
        if return_sparsity:
            return o, real_topk / nblk
        return o


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
