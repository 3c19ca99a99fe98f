"""f1(c): one denoise step of the UNMODIFIED reference WanModel (Wan2.1-T2V-1.3B configuration, random-init weights, quantised on
the GPU by the reference's own surgery) running on this repo's operators through install(), timed eagerly and as a CUDA graph,
next to turbodiffusion_b200.block.WanHotPath (the fused composition bench.py times) on the same shapes.

    python tools/ref_model_step.py [--layers 30] > gpurun_out/ref_model_step.json

Needs oracle/_ref/py (oracle/stage_ref_py.py).  MEASUREMENT TOOL: nothing here is imported by the product.
"""
import argparse
import json
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF_PY = os.path.join(ROOT, "oracle", "_ref", "py")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=30)
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    if not os.path.isdir(os.path.join(REF_PY, "rcm")):
        raise SystemExit("oracle/_ref/py is not staged")
    import turbodiffusion_b200
    turbodiffusion_b200.install()
    for p in (REF_PY, os.path.join(REF_PY, "inference")):
        sys.path.insert(0, p)
    stub = types.ModuleType("rcm.utils.model_utils")
    stub.load_state_dict = lambda *a, **k: {}
    sys.modules.setdefault("rcm.utils.model_utils", stub)
    import modify_model as mm

    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    m = mm.WanModel2pt1(dim=1536, eps=1e-6, ffn_dim=8960, freq_dim=256, in_dim=16, model_type="t2v", num_heads=12,
                        num_layers=args.layers, out_dim=16, text_len=512)
    m.init_weights()
    with torch.no_grad():
        for blk in m.blocks:
            blk.modulation.normal_(0, 0.02)
    m = m.to(torch.bfloat16)
    mm.replace_attention(m, "sagesla", 0.1)
    m = m.to(dev)
    mm.replace_linear_norm(m, replace_linear=True, replace_norm=True, quantize=True)
    m = m.to(dev).eval()

    x = torch.randn(1, 16, 21, 60, 104, device=dev).bfloat16()      # 81 frames at 832x480 -> latent 21 x 60 x 104 -> L = 32760
    t = torch.tensor([[500.0]], device=dev)
    ctx = torch.randn(1, 512, 4096, device=dev).bfloat16()

    def timed(fn, n):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(n):
            fn()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / n

    out = {"model": f"reference WanModel (rcm/networks/wan2pt1.py), dim 1536, 12 heads, ffn 8960, {args.layers} blocks, SageSLA top-k 0.1, "
                    "Int8Linear / FastNorm from turbodiffusion_b200 via install()", "L": 32760}
    with torch.no_grad():
        for _ in range(3):
            y = m(x, t, ctx)
        torch.cuda.synchronize()
        assert torch.isfinite(y.float()).all()
        out["reference_model_eager_ms"] = timed(lambda: m(x, t, ctx), args.steps)
        try:
            g = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                m(x, t, ctx)
            torch.cuda.current_stream().wait_stream(side)
            with torch.cuda.graph(g):
                yg = m(x, t, ctx)
            for _ in range(2):
                g.replay()
            out["reference_model_cuda_graph_ms"] = timed(g.replay, args.steps)
        except Exception as ex:  # noqa: BLE001
            out["reference_model_cuda_graph_error"] = f"{type(ex).__name__}: {str(ex)[:300]}"

    # the fused composition on the same shapes
    from turbodiffusion_b200.block import WanHotPath
    from turbodiffusion_b200.ops import wan_rope_angles
    del m
    torch.cuda.empty_cache()
    hp = WanHotPath(1536, 8960, 12, args.layers, dev, topk=0.1, seed=1)
    xs = torch.randn(32760, 1536, device=dev).bfloat16()
    e0 = torch.randn(6, 1536, device=dev) * 0.1
    cs = torch.randn(512, 1536, device=dev).bfloat16()
    ang = wan_rope_angles(21, 30, 52, 128).to(dev)
    for _ in range(3):
        hp.step(xs, e0, ang, cs)
    torch.cuda.synchronize()
    out["wan_hot_path_eager_ms"] = timed(lambda: hp.step(xs, e0, ang, cs), args.steps)
    g2 = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        hp.step(xs, e0, ang, cs)
    torch.cuda.current_stream().wait_stream(side)
    with torch.cuda.graph(g2):
        hp.step(xs, e0, ang, cs)
    for _ in range(2):
        g2.replay()
    out["wan_hot_path_cuda_graph_ms"] = timed(g2.replay, args.steps)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
