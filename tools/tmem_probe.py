"""TMEM read-throughput probe (see tdb200_selftest_tmem_read)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from turbodiffusion_b200._lib import lib, check, ptr, stream_ptr
dev = torch.device("cuda:0")
cyc = torch.zeros(256, dtype=torch.int64, device=dev)
sink = torch.zeros(4, device=dev)
iters = 2000
for convert in (0, 1):
    for warps in (4, 8, 12, 16):
        for _ in range(2):
            check(lib().tdb200_selftest_tmem_read(warps, iters, convert, ptr(cyc), ptr(sink), stream_ptr(dev)), "probe")
        torch.cuda.synchronize()
        c = cyc[:148].float().mean().item()
        bytes_ = warps * iters * 8192
        print(json.dumps({"convert": convert, "warps": warps, "cycles": c, "B_per_clk_per_SM": bytes_ / c,
                          "cycles_per_128KB": 131072 / (bytes_ / c)}))
