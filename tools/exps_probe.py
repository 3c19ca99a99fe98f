"""Duration of the fused attention's exponential pass (64 scores per thread) on register data, per variant and warps per scheduler
(tdb200_selftest_softmax_exps).  XU floor: 64 ex2 x 8 clk = 512 clk per warp and scheduler.

    python tools/exps_probe.py > gpurun_out/r02_exps_probe.json
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
names = ["as shipped", "no 16-bit packing", "no row sums", "no packing, no sums", "ex2 only"]
out = {}
for variant, name in enumerate(names):
    for warps in (4, 8):
        iters = 400
        check(lib().tdb200_selftest_softmax_exps(variant, warps, iters, ptr(cyc), ptr(sink), 0), "selftest_softmax_exps")
        torch.cuda.synchronize()
        out[f"{name}/warps_per_scheduler={warps // 4}"] = round(cyc.float().median().item() / iters, 1)
print(json.dumps(out, indent=1))
