"""Is the denoise step GPU-bound or host-launch-bound?  Times host enqueue vs GPU execution of one step."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from turbodiffusion_b200.block import WanHotPath
from turbodiffusion_b200.ops import wan_rope_angles
dev = torch.device("cuda:0")
L, dim, heads, ffn, layers = 32760, 1536, 12, 8960, 30
model = WanHotPath(dim, ffn, heads, layers, dev, topk=0.1, seed=1234)
g = torch.Generator().manual_seed(7)
x = torch.randn(L, dim, generator=g).bfloat16().to(dev)
e0 = (torch.randn(6, dim, generator=g) * 0.1).to(dev)
ctx = torch.randn(512, dim, generator=g).bfloat16().to(dev)
ang = wan_rope_angles(21, 30, 52, 128, dev)
for _ in range(2):
    model.step(x, e0, ang, ctx)
torch.cuda.synchronize()
res = {}
t0 = time.perf_counter(); s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
s.record(); model.step(x, e0, ang, ctx); e.record(); t1 = time.perf_counter()
torch.cuda.synchronize(); t2 = time.perf_counter()
res["host_enqueue_ms"] = (t1 - t0) * 1e3; res["gpu_ms"] = s.elapsed_time(e); res["wall_ms"] = (t2 - t0) * 1e3
# CUDA graph of the whole step
gr = torch.cuda.CUDAGraph()
sx = x.clone()
with torch.cuda.graph(gr):
    out = model.step(sx, e0, ang, ctx)
torch.cuda.synchronize()
for _ in range(2):
    gr.replay()
torch.cuda.synchronize()
s.record()
for _ in range(3):
    gr.replay()
e.record(); torch.cuda.synchronize()
res["graph_ms_per_step"] = s.elapsed_time(e) / 3
ref = model.step(x, e0, ang, ctx)
res["graph_equals_eager"] = bool((out.float() - ref.float()).abs().max().item() < 0.5)
print(json.dumps(res))
