"""Per-kernel timings on the GPU (CUDA events, L2 flushed between iterations).  Writes gpurun_out/microbench.jsonl.

    python tools/microbench.py [--filter gemm] [--iters 20]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import turbodiffusion_b200.ops as ops  # noqa: E402
from turbodiffusion_b200 import turbo_diffusion_ops as tdo  # noqa: E402

PEAKS = {"hbm_gbs": 6564.5, "bf16_tflops": 1736.7}
try:
    PEAKS.update(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))))
except Exception:
    pass

dev = torch.device("cuda:0")
_flush = None


def flush_l2():
    global _flush
    if _flush is None:
        _flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    _flush.zero_()


def timeit(fn, iters, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    times = []
    for _ in range(iters):
        flush_l2()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        times.append(s.elapsed_time(e))
    times.sort()
    return times[len(times) // 2], times[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--filter", default="")
    ap.add_argument("--iters", type=int, default=15)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "microbench.jsonl"))
    args = ap.parse_args()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    results = []
    filters = [f for f in args.filter.split(",") if f] or [""]

    def match(name):
        return any(f in name for f in filters)

    def record(name, ms_med, ms_min, flops=None, bytes_=None, **extra):
        r = {"name": name, "ms_median": round(ms_med, 4), "ms_min": round(ms_min, 4)}
        if flops:
            r["tflops"] = round(flops / ms_med / 1e9, 1)
            r["frac_bf16_peak"] = round(flops / ms_med / 1e9 / PEAKS["bf16_tflops"], 3)
        if bytes_:
            r["gbs"] = round(bytes_ / ms_med / 1e6, 1)
            r["frac_hbm_peak"] = round(bytes_ / ms_med / 1e6 / PEAKS["hbm_gbs"], 3)
        r.update(extra)
        results.append(r)
        print(json.dumps(r), flush=True)

    shapes = {"A": (32760, 1536, 8960), "B": (75600, 5120, 13824)}
    for tag, (L, dim, ffn) in shapes.items():
        gemms = [("qkv_fused", L, 3 * dim, dim), ("o_proj", L, dim, dim), ("ffn_up", L, ffn, dim), ("ffn_down", L, dim, ffn)]
        for name, m, n, k in gemms:
            full = f"gemm_w8a8/{tag}/{name}/{m}x{n}x{k}"
            if not match(full):
                continue
            a = torch.randint(-128, 128, (m, k), device=dev, dtype=torch.int8)
            b = torch.randint(-128, 128, (n, k), device=dev, dtype=torch.int8)
            a_s = torch.rand((m + 127) // 128, k // 128, device=dev) * 0.01
            b_s = torch.rand((n + 127) // 128, k // 128, device=dev) * 0.01
            c = torch.empty(m, n, dtype=torch.bfloat16, device=dev)
            med, mn = timeit(lambda: tdo.gemm_cuda(a, a_s, b, b_s, c), args.iters)
            record(full, med, mn, flops=2.0 * m * n * k, bytes_=m * k + n * k + 2 * m * n)
            bias = torch.randn(n, device=dev).bfloat16()
            if name in ("o_proj", "ffn_down"):   # the epilogues the block actually runs: +bias; +bias+GELU+int8 block quant
                med, mn = timeit(lambda: tdo.gemm_cuda_swizzle_bias(a, a_s, b, b_s, c, bias), args.iters)
                record(full.replace("gemm_w8a8/", "gemm_w8a8_bias/"), med, mn, flops=2.0 * m * n * k, bytes_=m * k + n * k + 2 * m * n)
            if name == "ffn_up":
                med, mn = timeit(lambda: tdo.gemm_cuda_quant_out(a, a_s, b, b_s, bias, torch.bfloat16, gelu=True), args.iters)
                record(full.replace("gemm_w8a8/", "gemm_w8a8_gelu_quant/"), med, mn, flops=2.0 * m * n * k, bytes_=m * k + n * k + m * n)
            del a, b, c
        for name, k in (("dim", dim), ("ffn", ffn)):
            full = f"quant_int8/{tag}/{name}/{L}x{k}"
            if match(full):
                x = torch.randn(L, k, device=dev, dtype=torch.bfloat16)
                q = torch.empty(L, k, dtype=torch.int8, device=dev)
                s = torch.empty((L + 127) // 128, k // 128, device=dev)
                med, mn = timeit(lambda: tdo.quant_cuda(x, q, s), args.iters)
                record(full, med, mn, bytes_=3 * L * k)
                del x, q
        full = f"gelu_quant_int8/{tag}/ffn/{L}x{ffn}"
        if match(full):
            xg = torch.randn(L, ffn, device=dev, dtype=torch.bfloat16)
            med, mn = timeit(lambda: tdo.gelu_quant_cuda(xg), args.iters)
            record(full, med, mn, bytes_=3 * L * ffn)
            del xg
        x = torch.randn(L, dim, device=dev, dtype=torch.bfloat16)
        w = torch.rand(dim, device=dev) + 0.5
        sc, sh = torch.randn(dim, device=dev) * 0.1, torch.randn(dim, device=dev) * 0.1
        rows = [
            (f"fast_layernorm/{tag}/{L}x{dim}", lambda: ops.fast_layernorm(x, None, None, 1e-6), 4 * L * dim),
            (f"fast_rmsnorm/{tag}/{L}x{dim}", lambda: ops.fast_rmsnorm(x, w, 1e-6), 4 * L * dim),
            (f"ln_modulate/{tag}/{L}x{dim}", lambda: ops.layernorm_modulate(x, sc, sh, 1e-6), 4 * L * dim),
            (f"ln_modulate_quant/{tag}/{L}x{dim}", lambda: ops.layernorm_modulate_quant(x, sc, sh, 1e-6), 3 * L * dim),
            (f"gate_residual/{tag}/{L}x{dim}", lambda: ops.gate_residual(x, x, sc), 6 * L * dim),
        ]
        heads = dim // 128
        ang = torch.rand(L, 64, device=dev) * 20
        rows.append((f"rmsnorm_rope/{tag}/{L}x{dim}", lambda: ops.rmsnorm_rope(x, w, ang, 1e-6, heads), 4 * L * dim))
        for full, fn, nbytes in rows:
            if match(full):
                med, mn = timeit(fn, args.iters)
                record(full, med, mn, bytes_=nbytes)
        del x

    # LTX-2 per-row post-scale W8A8 (no per-K-block dequant): M = 28672 video tokens, dim 4096, ffn 16384
    from turbodiffusion_b200 import ltx
    for name, m, n, k in [("ltx_proj", 28672, 4096, 4096), ("ltx_ffn_up", 28672, 16384, 4096), ("ltx_ffn_down", 28672, 4096, 16384)]:
        full = f"gemm_w8a8_rowwise/C/{name}/{m}x{n}x{k}"
        if not match(full):
            continue
        a = torch.randint(-128, 128, (m, k), device=dev, dtype=torch.int8)
        b = torch.randint(-128, 128, (n, k), device=dev, dtype=torch.int8)
        a_s, b_s = torch.rand(m, device=dev) * 0.01, torch.rand(n, device=dev) * 0.01
        bias = torch.randn(n, device=dev).bfloat16()
        med, mn = timeit(lambda: ltx.gemm_int8_post_scale_bias(a, a_s, b, b_s, bias), args.iters)
        record(full, med, mn, flops=2.0 * m * n * k, bytes_=m * k + n * k + 2 * m * n)
        del a, b

    # the SageSLA pipeline on synthetic q,k,v [1, L, H, D] bf16 (K with a per-channel bias so
    # smooth-K matters): preparation, block map, linear moments, the fused attention kernel alone, and the whole module.
    # Attention FLOPs = H * Mblk * topk_blocks * 4*128*64*D (QK^T + PV of the selected blocks).
    from turbodiffusion_b200.SLA import SageSparseLinearAttention
    from turbodiffusion_b200.SLA.core import attn_fwd, linear_moments
    from turbodiffusion_b200.SLA.utils import block_map_from_pools, quant_qk
    sla_cases = [("A", 32760, 12, 128, (0.1, 0.15, 1.0)), ("A64", 32760, 24, 64, (0.1,)), ("B", 75600, 40, 128, (0.1,)),
                 ("C", 28672, 32, 128, (0.3,))]
    for tag, L, H, D, ratios in sla_cases:
        prefixes = [f"sla_{part}/{tag}/" for part in ("prep", "block_map", "moments", "attn", "module")]
        if args.filter and not any(f in n or f.startswith(n) for n in prefixes for f in filters):
            continue  # (a selected case runs all of its parts: they share the prepared tensors)
        g = torch.Generator(device="cuda").manual_seed(0)
        q = torch.randn(1, L, H, D, device=dev, generator=g).bfloat16()
        k = (torch.randn(1, L, H, D, device=dev, generator=g) + 2 * torch.randn(1, 1, H, D, device=dev, generator=g)).bfloat16()
        v = torch.randn(1, L, H, D, device=dev, generator=g).bfloat16()
        mblk, nblk = (L + 127) // 128, (L + 63) // 64
        if D == 128:
            med, mn = timeit(lambda: quant_qk(q, k), args.iters)
            record(f"sla_prep/{tag}/{L}x{H}x{D}", med, mn, bytes_=L * H * D * (2 + 1) + L * H * D * (2 + 2 + 1))
            prep = quant_qk(q, k)
            med, mn = timeit(lambda: linear_moments(k, v), args.iters)
            record(f"sla_moments/{tag}/{L}x{H}x{D}", med, mn, flops=2.0 * L * D * D * H, bytes_=2 * L * H * D * 2)
            kv, ksum = linear_moments(k, v)
            kvw = kv.bfloat16().contiguous()
            pb = torch.zeros(D, device=dev)
            for r in ratios:
                topk = min(nblk, int(r * nblk))
                med, mn = timeit(lambda: block_map_from_pools(prep.q_pool, prep.k_pool, topk), args.iters)
                record(f"sla_block_map/{tag}/topk{r}", med, mn)
                _, lut = block_map_from_pools(prep.q_pool, prep.k_pool, topk)
                med, mn = timeit(lambda: attn_fwd(prep, v, q, lut, topk, kvw, ksum, pb, D ** -0.5), args.iters)
                record(f"sla_attn/{tag}/topk{r}/{topk}of{nblk}", med, mn, flops=float(H) * mblk * topk * 4 * 128 * 64 * D,
                       bytes_=4 * L * H * D * 2)
        for r in ratios:
            mod = SageSparseLinearAttention(D, r).to(dev)
            with torch.no_grad():
                mod.proj_l.weight.normal_(0, 0.05)
            med, mn = timeit(lambda: mod(q, k, v), max(3, args.iters // 2))
            topk = min(nblk, int(r * nblk))
            record(f"sla_module/{tag}/topk{r}", med, mn, flops=float(H) * mblk * topk * 4 * 128 * 64 * D + 6.0 * L * D * D * H)
        del q, k, v

    with open(args.out, "w") as f:
        for r in results:
            f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
