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
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "ref_vs_ours.json"), "w"), indent=1)
