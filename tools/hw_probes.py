"""Hardware probes behind the roofline denominators and two kernel-design decisions (run on the GPU box):

  * INT8 tensor peak, measured the way MEASURED_PEAKS.json measures bf16: torch._int_mm 8192^3 (cuBLASLt), best of 10
    (burst) and back to back for 4 s (sustained); cross-checked with this repo's per-row-scale W8A8 GEMM (tcgen05 kind::i8
    with one epilogue per tile) on the same shape.
  * XU throughput of ex2 / tanh in f32 and packed 16-bit forms (tdb200_selftest_mufu).
  * TMEM -> register read throughput with and without the dequant arithmetic (tdb200_selftest_tmem_read).

    python tools/hw_probes.py > gpurun_out/hw_probes.json
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from turbodiffusion_b200._lib import check, lib, ptr  # noqa: E402

dev = torch.device("cuda:0")
out = {"gpu": torch.cuda.get_device_name(0)}


def time_fn(fn, n):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


# ---- INT8 peak
N = 8192
a = torch.randint(-128, 128, (N, N), device=dev, dtype=torch.int8)
b = torch.randint(-128, 128, (N, N), device=dev, dtype=torch.int8)
flops = 2.0 * N ** 3
try:
    for _ in range(3):
        torch._int_mm(a, b.t())
    torch.cuda.synchronize()
    burst = min(time_fn(lambda: torch._int_mm(a, b.t()), 1) for _ in range(10))
    t0, n_s, ms_s = time.time(), 0, 0.0
    while time.time() - t0 < 4.0:
        ms_s += time_fn(lambda: torch._int_mm(a, b.t()), 20) * 20
        n_s += 20
    out["int8_tops_burst_cublaslt"] = flops / burst / 1e9
    out["int8_tops_sustained_cublaslt"] = flops / (ms_s / n_s) / 1e9
except Exception as ex:  # noqa: BLE001
    out["int8_cublaslt_error"] = f"{type(ex).__name__}: {ex}"
xa = torch.randn(N, N, device=dev).bfloat16()
xb = torch.randn(N, N, device=dev).bfloat16()
for _ in range(3):
    xa @ xb.t()
torch.cuda.synchronize()
out["bf16_tflops_burst_here"] = flops / min(time_fn(lambda: xa @ xb.t(), 1) for _ in range(10)) / 1e9
from turbodiffusion_b200 import ltx  # noqa: E402
a_s, b_s = torch.rand(N, device=dev) * 0.01, torch.rand(N, device=dev) * 0.01
bias = torch.zeros(N, device=dev).bfloat16()
for _ in range(3):
    ltx.gemm_int8_post_scale_bias(a, a_s, b, b_s, bias)
torch.cuda.synchronize()
out["int8_tops_burst_tdb200_rowwise"] = flops / min(time_fn(lambda: ltx.gemm_int8_post_scale_bias(a, a_s, b, b_s, bias), 1) for _ in range(10)) / 1e9
t0, n_s, ms_s = time.time(), 0, 0.0
while time.time() - t0 < 4.0:
    ms_s += time_fn(lambda: ltx.gemm_int8_post_scale_bias(a, a_s, b, b_s, bias), 20) * 20
    n_s += 20
out["int8_tops_sustained_tdb200_rowwise"] = flops / (ms_s / n_s) / 1e9
out["int8_how"] = ("torch._int_mm 8192^3 (2*N^3 ops): best of 10 single launches (burst) and back to back for 4 s (sustained), "
                   "CUDA events; same protocol as MEASURED_PEAKS.json's bf16 entries; rowwise = tdb200_gemm_w8a8_rowwise")

# ---- MUFU forms
sms = torch.cuda.get_device_properties(0).multi_processor_count
cyc = torch.zeros(sms, dtype=torch.int64, device=dev)
sink = torch.zeros(4, device=dev)
names = ["ex2.f32", "ex2.bf16x2", "ex2.f16x2", "tanh.f32", "tanh.bf16x2", "tanh.f16x2"]
mufu = {}
for mode, name in enumerate(names):
    for warps in (8, 16):
        iters = 2000
        check(lib().tdb200_selftest_mufu(mode, warps, iters, ptr(cyc), ptr(sink), 0), "selftest_mufu")
        torch.cuda.synchronize()
        c = cyc.float().median().item()
        per_clk = warps * 32 * 8 * iters * (1 if name.endswith("f32") else 2) / c
        mufu[f"{name}/warps{warps}"] = {"cycles": c, "results_per_clk_per_sm": round(per_clk, 2)}
out["mufu"] = mufu

# ---- TMEM read
tm = {}
for warps in (4, 8, 16):
    for convert in (0, 1):
        iters = 512
        check(lib().tdb200_selftest_tmem_read(warps, iters, convert, ptr(cyc), ptr(sink), 0), "selftest_tmem_read")
        torch.cuda.synchronize()
        c = cyc.float().median().item()
        tm[f"warps{warps}/convert{convert}"] = {"cycles": c, "bytes_per_clk_per_sm": round(warps * 8192 * iters / c, 1),
                                                "clk_per_128KB_kblock": round(c / (warps * 8192 * iters) * 131072, 1)}
out["tmem_read"] = tm
print(json.dumps(out, indent=1))
