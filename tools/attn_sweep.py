"""time(attn_fwd) vs number of selected key blocks: t = a + b*T separates per-CTA overhead from per-iteration cost."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from turbodiffusion_b200.SLA.core import attn_fwd, linear_moments
from turbodiffusion_b200.SLA.utils import quant_qk, block_map_from_pools
dev = torch.device("cuda:0")
L, H, D = 32760, 12, 128
g = torch.Generator(device="cuda").manual_seed(0)
q = torch.randn(1, L, H, D, device=dev, generator=g).bfloat16()
k = (torch.randn(1, L, H, D, device=dev, generator=g) + torch.randn(1, 1, H, D, device=dev, generator=g)).bfloat16()
v = torch.randn(1, L, H, D, device=dev, generator=g).bfloat16()
prep = quant_qk(q, k)
kv, ksum = linear_moments(k, v)
kvw = (torch.eye(D, device=dev)[None, None] * 0 + kv).bfloat16().contiguous()
pb = torch.zeros(D, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for topk in (4, 13, 26, 51, 102, 204):
    _, lut = block_map_from_pools(prep.q_pool, prep.k_pool, topk)
    ts = []
    for it in range(6):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); attn_fwd(prep, v, q, lut, topk, kvw, ksum, pb, D ** -0.5); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    ctas = 256 * H
    us = ts[len(ts) // 2] * 1e3
    print(json.dumps({"topk_blocks": topk, "us": round(us, 1), "us_per_cta_slot": round(us / (ctas / 296), 2),
                      "ns_per_iter_per_cta": round(us * 1e3 / (ctas / 296) / topk, 1)}))

# ---- phase trace of one softmax warp (see g_attn_trace in sla_attn.cu)
from turbodiffusion_b200._lib import lib, check, ptr
trace = torch.zeros(2 * 64 * 8, dtype=torch.int64, device=dev)
check(lib().tdb200_debug_set_attn_trace(ptr(trace)), "set trace")
_, lut = block_map_from_pools(prep.q_pool, prep.k_pool, 51)
attn_fwd(prep, v, q, lut, 51, kvw, ksum, pb, D ** -0.5)
torch.cuda.synchronize()
check(lib().tdb200_debug_set_attn_trace(None), "clear trace")
t = trace.cpu().view(2, 64, 8)
names = ["wait_S", "ldtm", "max+vote", "exps", "wait_Pfree", "store+publish", "->next"]
for slot in range(2):
    tt = t[slot, :51]
    d = torch.stack([tt[:, k + 1] - tt[:, k] for k in range(6)] + [torch.cat([tt[1:, 0] - tt[:-1, 6], torch.zeros(1, dtype=torch.int64)])], 1).float()
    per_iter = (tt[1:, 0] - tt[:-1, 0]).float()
    print(json.dumps({"slot": slot, "iter_cycles_median": per_iter.median().item(), "iter_cycles_mean": per_iter.mean().item(),
                      "phase_median": {n: d[5:45, i].median().item() for i, n in enumerate(names)},
                      "wait_S_next+ld_issue_median": (tt[5:45, 7] - tt[5:45, 4]).float().median().item(),
                      "wait_PV_prev_median": (tt[5:45, 5] - tt[5:45, 7]).float().median().item(),
                      "loop_total": int(t[slot, 63, 7] - tt[0, 0]),
                      "tile_marks_clk_from_entry": {n: int(t[slot, 62, i] - t[slot, 62, 0]) for i, n in
                                                    enumerate(["entry", "first_S", "loop_exit", "phi_published", "OL_ready", "stores_issued"])}}))
