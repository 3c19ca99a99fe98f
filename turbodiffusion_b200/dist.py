"""Sequence (video-token) parallelism for the hot path: one process per GPU, torch.distributed for the plumbing.

Every GEMM / norm / modulation / RoPE / FFN in the block is row-independent, so rank r simply owns the token rows
[row_begin, row_end).  Only self-attention needs an exchange (SURVEY 8e / BASELINE north star):

    1. all-gather of the local K and V slabs (16-bit, [rows, H, D]) into one [L_pad, H, D] slab per tensor;
    2. all-reduce (sum) of the linear-attention moments  phi(K)^T V  [H,D,D]  and  sum phi(K)  [H,D], which
       tdb200_sla_linear_moments accumulates over the LOCAL rows only;
    3. everything else (key mean, INT8 K, block map over all key blocks, fused attention over the gathered K/V) is computed
       locally for this rank's query rows.

Shard boundaries are multiples of 128 rows, so 128x128 quant blocks, 128-row query blocks and 64-row key blocks never
straddle ranks and the result equals the single-GPU computation block for block.  Only the last rank may be short; the
all-gather pads it with zero rows at the very END of the sequence, so the gathered slab is simply the full tensor followed
by < 128*world padding rows that the kernels never read (lk = L).

The reference's own scheme is Ulysses all-to-all (rcm/utils/a2a_cp.py:66-182; heads <-> sequence), which needs
H % world == 0 (12 heads do not split 8 ways); it is not used by the inference scripts.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist


def cdiv(a: int, b: int) -> int:
    return (a + b - 1) // b


def shard_rows(total_rows: int, world: int, rank: int, align: int = 128):
    """128-aligned contiguous shards; returns (row_begin, row_end, rows_per_rank_padded)."""
    blocks = cdiv(total_rows, align)
    per_rank = cdiv(blocks, world) * align
    begin = min(total_rows, rank * per_rank)
    end = min(total_rows, (rank + 1) * per_rank)
    return begin, end, per_rank


class GpuPrimitives:
    """The C-ABI kernels (default).  The CPU test injects oracle implementations with the same signatures."""

    def __init__(self, sla_module):
        self.sla = sla_module

    def prepare_q(self, q):
        from .SLA.utils import quant_q_only
        return quant_q_only(q)

    def attention(self, q, k_full, v_full, lk, kv, ksum, qprep=None):
        from .SLA.core import attn_fwd
        from .SLA.utils import block_map_from_pools, quant_k_into, quant_q_only
        sla = self.sla
        d = q.shape[-1]
        prep = quant_k_into(qprep if qprep is not None else quant_q_only(q), k_full, lk)
        topk = min(prep.nblk, int(sla.topk * prep.nblk))
        _, lut = block_map_from_pools(prep.q_pool, prep.k_pool, topk)
        kvw = torch.matmul(sla.proj_l.weight.float(), kv).to(q.dtype).contiguous()
        return attn_fwd(prep, v_full, q, lut, topk, kvw, ksum, sla.proj_l.bias.float().contiguous(), d ** -0.5, lk=lk)

    def moments(self, k_local, v_local):
        from .SLA.core import linear_moments
        return linear_moments(k_local, v_local)


class SPAttention:
    """Drop-in for the block's attention callable: (q, k, v) local [1, rows, H, D] -> [1, rows, H, D].

    The block calls start_kv(k, v) as soon as K and V exist: both all-gathers and the moment all-reduces are issued
    asynchronously (NCCL stream) and overlap with the Q projection, Q RMSNorm+RoPE and the Q-side quantisation that the
    block and __call__ run before waiting on them."""

    def __init__(self, sp: "SequenceParallel", prims):
        self.sp, self.prims = sp, prims
        self._bufs = {}
        self._pending = None
        self._pending_k = None

    def _gather_async(self, name: str, t: torch.Tensor):
        """t [1, rows, H, D] -> ([1, world*rows_pad, H, D] buffer, NCCL work handle); zero-padded tail of the last rank."""
        sp = self.sp
        _, rows, h, d = t.shape
        key = (name, h, d, t.dtype, t.device)
        if key not in self._bufs:
            self._bufs[key] = (torch.zeros(sp.rows_pad, h, d, dtype=t.dtype, device=t.device),
                               torch.zeros(sp.world * sp.rows_pad, h, d, dtype=t.dtype, device=t.device))
        send, recv = self._bufs[key]
        send[:rows].copy_(t[0])
        work = dist.all_gather_into_tensor(recv, send, group=sp.group, async_op=True)
        return recv.unsqueeze(0), work

    def _cdt(self, t):
        return self.prims.sla.dtype if hasattr(self.prims, "sla") else t.dtype

    def start_k(self, k):
        """K exists (projected, normalised, rotated): start its all-gather; the V and Q projections run under it."""
        k = k.to(self._cdt(k)).contiguous()
        k_full, wk = self._gather_async("k", k)
        self._pending_k = (k, k_full, wk)

    def start_kv(self, k, v):
        sp = self.sp
        cdt = self._cdt(k)
        if getattr(self, "_pending_k", None) is None:
            self.start_k(k)
        k, k_full, wk = self._pending_k
        self._pending_k = None
        v = v.to(cdt).contiguous()
        v_full, wv = self._gather_async("v", v)
        kv, ksum = self.prims.moments(k, v)
        w1 = dist.all_reduce(kv, group=sp.group, async_op=True)
        w2 = dist.all_reduce(ksum, group=sp.group, async_op=True)
        self._pending = (k_full, v_full, kv, ksum, (wk, wv, w1, w2))

    def __call__(self, q, k, v):
        dtype = q.dtype
        q = q.to(self._cdt(q)).contiguous()
        if self._pending is None:
            self.start_kv(k, v)
        k_full, v_full, kv, ksum, works = self._pending
        self._pending = None
        qprep = self.prims.prepare_q(q) if hasattr(self.prims, "prepare_q") else None  # overlaps with the collectives
        for w in works:
            w.wait()
        out = self.prims.attention(q, k_full, v_full, self.sp.total_rows, kv, ksum, qprep)
        return out.to(dtype)


class SequenceParallel:
    def __init__(self, total_rows: int, world: Optional[int] = None, rank: Optional[int] = None, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if world is None else world
        self.rank = dist.get_rank(group) if rank is None else rank
        self.total_rows = total_rows
        self.row_begin, self.row_end, self.rows_pad = shard_rows(total_rows, self.world, self.rank)
        self.local_rows = self.row_end - self.row_begin
        if self.local_rows <= 0:
            raise ValueError(f"rank {self.rank} owns no rows: L={total_rows} is too short for {self.world} ranks of 128-row blocks")

    def install(self, model) -> None:
        """Replace every block's attention callable by the sequence-parallel one (the reference seam is
        `WanSelfAttention.attn_op.local_attn`, inference/modify_model.py:48-52)."""
        for blk in model.blocks:
            blk.attn_hook = SPAttention(self, GpuPrimitives(blk.sla))

    def scatter(self, full: torch.Tensor) -> torch.Tensor:
        return full[self.row_begin:self.row_end].contiguous()

    def gather_rows(self, local: torch.Tensor) -> torch.Tensor:
        """All-gather of the step output rows (context_parallel.py:60-91 equivalent); local [rows, C] -> [L, C]."""
        send = torch.zeros(self.rows_pad, *local.shape[1:], dtype=local.dtype, device=local.device)
        send[: local.shape[0]].copy_(local)
        recv = torch.empty(self.world * self.rows_pad, *local.shape[1:], dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(recv, send, group=self.group)
        return recv[: self.total_rows]
