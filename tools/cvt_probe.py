"""Which pipe do the 16-bit packing conversions use?  Instruction-issue rates per SM (tdb200_selftest_mufu modes 0, 6-10):
results per clk per SM of ex2, cvt.rn.bf16x2.f32 (F2FP), cvt.rn.f16x2.f32, F2I+I2F, FMUL, and an ex2/F2FP mix.

    python tools/cvt_probe.py > gpurun_out/r02_cvt_probe.json
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from turbodiffusion_b200._lib import check, lib, ptr  # noqa: E402

dev = torch.device("cuda:0")
sms = torch.cuda.get_device_properties(0).multi_processor_count
cyc = torch.zeros(sms, dtype=torch.int64, device=dev)
sink = torch.zeros(4, device=dev)
# (mode, name, instructions of interest per chain step)
cases = [(0, "ex2.approx.ftz.f32", 1), (6, "cvt.rn.bf16x2.f32 (+shl)", 1), (7, "cvt.rn.f16x2.f32 (+shl)", 1),
         (8, "cvt.rni.s32.f32 + cvt.rn.f32.s32", 2), (9, "mul.f32", 1), (10, "4 chains ex2 + 4 chains cvt.rn.bf16x2.f32", 1)]
out = {}
for mode, name, per in cases:
    for warps in (4, 8, 16):
        iters = 2000
        check(lib().tdb200_selftest_mufu(mode, warps, iters, ptr(cyc), ptr(sink), 0), "selftest_mufu")
        torch.cuda.synchronize()
        c = cyc.float().median().item()
        out[f"{name}/warps{warps}"] = {"cycles": c, "thread_instr_per_clk_per_sm": round(warps * 32 * 8 * iters * per / c, 2)}
print(json.dumps(out, indent=1))
