"""Head-to-head on the same B200: the reference's own CUDA kernels (oracle/_ref/turbo_diffusion_ops.so, built from the
unmodified sources; legacy mma.sync IMMA + cp.async) vs libtdb200 at the Wan shapes.  TEST/MEASUREMENT TOOL ONLY."""
import glob, importlib.util, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from turbodiffusion_b200 import turbo_diffusion_ops as ours

found = glob.glob(os.path.join(ROOT, "oracle", "_ref", "turbo_diffusion_ops*.so"))
if not found:
    raise SystemExit("oracle/_ref holds no build of the reference extension")
spec = importlib.util.spec_from_file_location("turbo_diffusion_ops", found[0])
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=8):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


out = []
for tag, (L, dim, ffn) in {"A": (32760, 1536, 8960), "B": (75600, 5120, 13824)}.items():
    for name, m, n, k in [("qkv/o/cross proj", L, dim, dim), ("ffn_up", L, ffn, dim), ("ffn_down", L, dim, ffn)]:
        a = torch.randint(-128, 128, (m, k), device=dev, dtype=torch.int8)
        b = torch.randint(-128, 128, (n, k), device=dev, dtype=torch.int8)
        a_s = torch.rand((m + 127) // 128, k // 128, device=dev) * 0.01
        b_s = torch.rand((n + 127) // 128, k // 128, device=dev) * 0.01
        c1 = torch.zeros(m, n, dtype=torch.bfloat16, device=dev)
        c2 = torch.zeros(m, n, dtype=torch.bfloat16, device=dev)
        t_ref = timeit(lambda: ref.gemm_cuda(a, a_s, b, b_s, c1))
        t_our = timeit(lambda: ours.gemm_cuda(a, a_s, b, b_s, c2))
        r = {"op": "gemm_w8a8", "shape": f"{tag} {name} {m}x{n}x{k}", "ref_ms": round(t_ref, 4), "ours_ms": round(t_our, 4),
             "speedup": round(t_ref / t_our, 2), "ref_tflops": round(2.0 * m * n * k / t_ref / 1e9, 1),
             "ours_tflops": round(2.0 * m * n * k / t_our / 1e9, 1), "bit_identical": bool(torch.equal(c1, c2))}
        out.append(r); print(json.dumps(r), flush=True)
        del a, b, c1, c2
    x = torch.randn(L, dim, device=dev, dtype=torch.bfloat16)
    q1 = torch.empty(L, dim, dtype=torch.int8, device=dev); s1 = torch.empty((L + 127) // 128, dim // 128, device=dev)
    q2, s2 = torch.empty_like(q1), torch.empty_like(s1)
    t_ref = timeit(lambda: ref.quant_cuda(x, q1, s1))
    t_our = timeit(lambda: ours.quant_cuda(x, q2, s2))
    r = {"op": "quant_int8", "shape": f"{tag} {L}x{dim}", "ref_ms": round(t_ref, 4), "ours_ms": round(t_our, 4),
         "speedup": round(t_ref / t_our, 2)}
    out.append(r); print(json.dumps(r), flush=True)

# ---- the reference's Triton kernels (SparseLinearAttention = SLA/kernel.py _attn_fwd + SLA/utils.py, FastRMSNorm / FastLayerNorm
#      = ops/core.py), run from the staged python sources (oracle/stage_ref_py.py) on the same tensors as this repo's modules.
#      SageSparseLinearAttention's own CUDA path needs the third-party SpargeAttn package: unobtainable here.
REF_PY = os.path.join(ROOT, "oracle", "_ref", "py")
if os.path.isdir(os.path.join(REF_PY, "SLA")):
    sys.modules["turbo_diffusion_ops"] = ref           # the reference `ops` package imports its compiled extension by name
    sys.path.insert(0, REF_PY)
    import importlib
    ref_sla = importlib.import_module("SLA")
    ref_ops = importlib.import_module("ops")
    import turbodiffusion_b200.SLA as our_sla
    import turbodiffusion_b200.ops as our_ops
    from oracle import td_oracle as O
    for tag, (L, H, D, ratio) in {"A 12x128": (32760, 12, 128, 0.1), "A 24x64": (32760, 24, 64, 0.1), "C 32x128": (28672, 32, 128, 0.3)}.items():
        g = torch.Generator(device="cuda").manual_seed(0)
        q = torch.randn(1, L, H, D, device=dev, generator=g).bfloat16()
        k = (torch.randn(1, L, H, D, device=dev, generator=g) + 2 * torch.randn(1, 1, H, D, device=dev, generator=g)).bfloat16()
        v = torch.randn(1, L, H, D, device=dev, generator=g).bfloat16()
        rm = ref_sla.SparseLinearAttention(D, ratio, BLKQ=128, BLKK=64).to(dev)
        om = our_sla.SageSparseLinearAttention(D, ratio).to(dev)
        with torch.no_grad():
            rm.proj_l.weight.normal_(0, 0.05)
            om.proj_l.weight.copy_(rm.proj_l.weight)
        try:
            with torch.no_grad():
                o_ref = rm(q, k, v)
                o_our = om(q, k, v)
            t_ref = timeit(lambda: rm(q, k, v), iters=5)
            t_our = timeit(lambda: om(q, k, v), iters=5)
            st = O.stats(o_our.float().cpu(), o_ref.float().cpu())
            r = {"op": "SLA module forward (reference: Triton SparseLinearAttention bf16; ours: SageSLA INT8 fused)",
                 "shape": f"{tag} q,k,v [1,{L},{H},{D}] topk {ratio}", "ref_ms": round(t_ref, 3), "ours_ms": round(t_our, 3),
                 "speedup": round(t_ref / t_our, 2), "rel_l2_vs_ref": st["rel_l2"], "cos_vs_ref": st["cos"]}
        except Exception as ex:  # noqa: BLE001
            r = {"op": "SLA module forward", "shape": tag, "error": f"{type(ex).__name__}: {str(ex)[:200]}"}
        out.append(r); print(json.dumps(r), flush=True)
        del q, k, v
    for tag, (L, dim) in {"A": (32760, 1536), "B": (75600, 5120)}.items():
        x = torch.randn(L, dim, device=dev, dtype=torch.bfloat16)
        for kind in ("rms", "ln"):
            if kind == "rms":
                rmod, omod = ref_ops.FastRMSNorm(dim, 1e-6).to(dev), our_ops.FastRMSNorm(dim, 1e-6).to(dev)
            else:
                rmod, omod = ref_ops.FastLayerNorm(dim, 1e-6).to(dev), our_ops.FastLayerNorm(dim, 1e-6).to(dev)
            try:
                y_ref, y_our = rmod(x[None]), omod(x[None])
                t_ref, t_our = timeit(lambda: rmod(x[None])), timeit(lambda: omod(x[None]))
                ulp = (y_ref.view(torch.int16).int() - y_our.view(torch.int16).int()).abs()
                r = {"op": f"Fast{'RMS' if kind == 'rms' else 'Layer'}Norm.forward (reference: x.float() + Triton kernel + cast)",
                     "shape": f"{tag} [1,{L},{dim}] bf16", "ref_ms": round(t_ref, 4), "ours_ms": round(t_our, 4),
                     "speedup": round(t_ref / t_our, 2), "max_ulp_diff": int(ulp.max()), "frac_differing": float((ulp > 0).float().mean())}
            except Exception as ex:  # noqa: BLE001
                r = {"op": f"Fast{kind}", "shape": tag, "error": f"{type(ex).__name__}: {str(ex)[:200]}"}
            out.append(r); print(json.dumps(r), flush=True)
        del x
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "ref_vs_ours.json"), "w"), indent=1)
