"""Benchmark of the Wan DiT denoise-step hot path (BASELINE.json metric: denoise-step latency & frames/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--shape A|B] [--impl b200|reference]
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One "step" = one pass of the hot path over one batch of synthetic input = one denoise step of the DiT block stack
(30 blocks for Wan2.1-T2V-1.3B at 480p: L = 32760 video tokens, dim 1536, 12 heads x 128, ffn 8960, 512 text tokens;
SageSLA top-k 0.1, W8A8 linears, FastNorm) = the loop at rcm/networks/wan2pt1.py:697-698 with random-init weights.
value = 81 frames / (4 denoise steps x step latency)  [frames/s], higher is better; ms_per_step is the step latency.

Multi-GPU (N > 1): the video-token axis is sharded over ranks (128-row aligned); every block does one NCCL all-gather
of the local K/V slab and one all-reduce of the linear-attention moments (turbodiffusion_b200/dist.py), "strong" scaling.

--impl reference times the reference's own CPU-runnable path (dense torch SDPA attention + the block's linears, the
oracle port) on the host cores, on a bounded sample scaled to the same metric.
"""
from __future__ import annotations

import argparse
import contextlib
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SHAPES = {
    # name: (L, dim, heads, ffn, layers, text_len, latent T,H,W after patching)
    "A": dict(name="Wan2.1-T2V-1.3B-480p", L=32760, dim=1536, heads=12, ffn=8960, layers=30, text=512, thw=(21, 30, 52)),
    "B": dict(name="Wan2.1-T2V-14B-720p", L=75600, dim=5120, heads=40, ffn=13824, layers=40, text=512, thw=(21, 45, 80)),
}
FRAMES, DENOISE_STEPS = 81, 4


def peaks():
    p = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "_source": "fallback"}
    try:
        p.update(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))))
        p["_source"] = "measured"
    except Exception:
        pass
    return p


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass

    def summary(self):
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        reasons = []
        for i, name in enumerate(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]):
            if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        mx = max(int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit())
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": mx, "reasons": reasons, "samples": len(sm)}


class GemmTimer:
    """CUDA-event timing of every W8A8 GEMM launch inside the timed region (same stream as the launch)."""

    def __init__(self):
        self.records = []

    @contextlib.contextmanager
    def __call__(self, m, n, k):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        yield
        e.record()
        self.records.append((s, e, 2.0 * m * n * k))

    def result(self):
        ms = sum(s.elapsed_time(e) for s, e, _ in self.records)
        fl = sum(f for _, _, f in self.records)
        return len(self.records), ms, fl


# ------------------------------------------------------------------------------------------------ CPU reference arm
def cpu_reference_sample(shape, threads=None):
    """The reference's CPU-runnable path on a bounded sample: dense F.scaled_dot_product_attention
    (rcm/utils/attention.py:152-166 body) + the block's linears (oracle port: bf16 matmuls of the same shapes), one block
    at L_s rows, scaled to a full denoise step: attention x (L/L_s)^2, linears x (L/L_s), x layers."""
    from oracle import td_oracle as O
    torch.set_num_threads(threads or os.cpu_count())
    cores = torch.get_num_threads()
    L, dim, heads, ffn, layers = shape["L"], shape["dim"], shape["heads"], shape["ffn"], shape["layers"]
    d = dim // heads
    ls = 2048
    g = torch.Generator().manual_seed(0)
    q, k, v = (torch.randn(1, heads, ls, d, generator=g).bfloat16() for _ in range(3))
    x = torch.randn(ls, dim, generator=g).bfloat16()
    w_sq = torch.randn(dim, dim, generator=g).bfloat16()
    w_up = torch.randn(ffn, dim, generator=g).bfloat16()
    w_dn = torch.randn(dim, ffn, generator=g).bfloat16()

    def one():
        t0 = time.perf_counter()
        O.dense_attention(q, k, v)
        t_attn = time.perf_counter() - t0
        t0 = time.perf_counter()
        for _ in range(6):  # self q,k,v,o + cross q,o  (cross k,v act on 512 text tokens: negligible)
            x @ w_sq.t()
        u = x @ w_up.t()
        torch.nn.functional.gelu(u, approximate="tanh") @ w_dn.t()
        t_lin = time.perf_counter() - t0
        return t_attn, t_lin

    one()
    t_attn, t_lin = one()
    step_s = layers * (t_attn * (L / ls) ** 2 + t_lin * (L / ls))
    return {"step_s": step_s, "cores": cores, "t_attn_sample_s": t_attn, "t_lin_sample_s": t_lin,
            "sample": f"1 block at L_s={ls} rows (dense SDPA [1,{heads},{ls},{d}] bf16 + 8 bf16 linears), scaled "
                      f"attention x(L/L_s)^2, linears x(L/L_s), x{layers} blocks"}


def run_reference(args, shape, rank, world):
    if rank != 0:
        return
    t0 = time.perf_counter()
    samples = []
    for i in range(args.warmup + args.steps):
        r = cpu_reference_sample(shape)
        if i >= args.warmup:
            samples.append(r)
        if time.perf_counter() - t0 > 240:
            break
    step_s = sum(r["step_s"] for r in samples) / len(samples)
    val = FRAMES / (DENOISE_STEPS * step_s)
    line = {"impl": "reference", "metric": "frames_per_sec_4step_81f", "value": val, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": len(samples), "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{shape['name']} DiT denoise step, reference CPU path (dense SDPA + linears)",
                       "L": shape["L"], "dim": shape["dim"], "heads": shape["heads"], "ffn": shape["ffn"], "layers": shape["layers"]},
            "cpu_baseline": {"value": val, "unit": "frames/s", "cores": samples[0]["cores"], "kind": "port",
                             "sample": samples[0]["sample"]},
            "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ B200 arm
def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE gemm_w8a8 launch from the committed `ncu --set full` summary
    (profiles/r01_ncu_gemm_final.txt: the ffn2 GEMM, M=32760 K=8960 N=1536, algorithmic bytes 293.5+13.8+100.6 = 407.9 MB).
    The roofline above is tensor-bound; traffic ~= algorithmic bytes shows no operand is re-read from HBM."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r01_ncu_gemm_final.txt")
    try:
        tot = 0.0
        for ln in open(path):
            for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                if ln.strip().startswith(key + " ="):
                    val, unit = ln.split("=")[1].split()[:2]
                    tot += float(val) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
        return {"traffic": tot or None, "traffic_algorithmic": 407.9e6,
                "traffic_source": "profiles/r01_ncu_gemm_final.txt (ffn2 GEMM 32760x8960x1536, one launch)"}
    except OSError:
        return {"traffic": None}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--shape", default="A", choices=list(SHAPES))
    ap.add_argument("--layers", type=int, default=None, help="override the number of blocks (diagnostics only)")
    ap.add_argument("--topk", type=float, default=0.1)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--sp-mode", default="auto", choices=["auto", "allgather", "ulysses"],
                    help="N>1 attention exchange: K/V all-gather + moment all-reduce, or head<->sequence all-to-all "
                         "(needs heads %% N == 0); auto = the mode that measured faster (dist.py pick_mode)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile", action="store_true", help="device-resident region only (for runs under ncu)")
    ap.add_argument("--no-graph", action="store_true", help="time the eager (Python-launched) step instead of the CUDA graph")
    args = ap.parse_args()
    if args.profile:
        args.no_graph = True
    shape = dict(SHAPES[args.shape])
    if args.layers:
        shape["layers"] = args.layers

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        return run_reference(args, shape, rank, world)

    import torch.distributed as dist
    from turbodiffusion_b200 import _lib
    from turbodiffusion_b200 import turbo_diffusion_ops as tdo
    from turbodiffusion_b200.block import WanHotPath
    _lib.lib()  # fail loudly when the CUDA library is missing
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    L, dim, heads, ffn, layers, text = (shape[k] for k in ("L", "dim", "heads", "ffn", "layers", "text"))
    d = dim // heads
    model = WanHotPath(dim, ffn, heads, layers, dev, topk=args.topk, seed=1234)
    if world > 1:
        from turbodiffusion_b200.dist import SequenceParallel
        sp = SequenceParallel(L, world, rank)
        sp_mode = sp.install(model, args.sp_mode)
        rows = sp.local_rows
        row0 = sp.row_begin
    else:
        rows, row0 = L, 0

    from turbodiffusion_b200.ops import wan_rope_angles
    g = torch.Generator().manual_seed(7)
    x_host = torch.randn(L, dim, generator=g).bfloat16()[row0:row0 + rows].contiguous().pin_memory()
    e0_host = (torch.randn(6, dim, generator=g) * 0.1).pin_memory()
    ctx_host = torch.randn(text, dim, generator=g).bfloat16().pin_memory()
    angles = wan_rope_angles(*shape["thw"], d)[row0:row0 + rows].contiguous().to(dev)
    out_host = torch.empty(rows, dim, dtype=torch.bfloat16).pin_memory()
    x, e0, ctx = x_host.to(dev), e0_host.to(dev), ctx_host.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    # ---------------- device-resident timing (inputs already in HBM)
    # (1) eager pass: every kernel launched from Python through the C ABI, W8A8 GEMM launches bracketed by CUDA events
    for _ in range(args.warmup):
        model.step(x, e0, angles, ctx)
    barrier()
    timer = GemmTimer()
    tdo.GEMM_TIMER = timer
    launches0 = _lib.LAUNCHES
    with ClockSampler(local_rank) as clk:
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(args.steps):
            y = model.step(x, e0, angles, ctx)
        e.record()
        barrier()
    tdo.GEMM_TIMER = None
    launches = _lib.LAUNCHES - launches0
    ms_eager = max_over_ranks(s.elapsed_time(e) / args.steps)
    n_gemm, gemm_ms, gemm_flops = timer.result()
    gemm_share = gemm_ms / (s.elapsed_time(e))

    # (2) the same step captured once into a CUDA graph (kernels + NCCL collectives) and replayed: removes the ~1000 host
    #     launches per step from the critical path, which matters once the per-rank GPU time shrinks (N > 1)
    graph, graph_note = None, "disabled"
    xs, es, cs = x.clone(), e0.clone(), ctx.clone()
    if not args.no_graph:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                model.step(xs, es, angles, cs)
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                ys = model.step(xs, es, angles, cs)
            graph_note = "cuda graph replay"
        except Exception as ex:  # noqa: BLE001
            graph, graph_note = None, f"capture failed ({type(ex).__name__}), eager timing reported"
    if graph is not None:
        for _ in range(args.warmup):
            graph.replay()
        with ClockSampler(local_rank) as clk:
            barrier()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(args.steps):
                graph.replay()
            e.record()
            barrier()
        ms_step = max_over_ranks(s.elapsed_time(e) / args.steps)
    else:
        ms_step = ms_eager

    if args.profile:
        if rank == 0:
            print(json.dumps({"profile_run": True, "ms_per_step": ms_step, "gpu_launches": launches}), flush=True)
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier()
            sys.stdout.flush()
            os._exit(0)
        return
    # ---------------- end to end through the public call with HOST buffers (H2D of the inputs + D2H of the result per step)
    def step_host():
        if graph is not None:  # static graph inputs are refilled from pinned host memory, result read back
            xs.copy_(x_host, non_blocking=True)
            es.copy_(e0_host, non_blocking=True)
            cs.copy_(ctx_host, non_blocking=True)
            graph.replay()
            out_host.copy_(ys, non_blocking=True)
            return
        xd = x_host.to(dev, non_blocking=True)
        ed = e0_host.to(dev, non_blocking=True)
        cd = ctx_host.to(dev, non_blocking=True)
        out_host.copy_(model.step(xd, ed, angles, cd), non_blocking=True)

    step_host()
    barrier()
    s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s2.record()
    for _ in range(args.steps):
        step_host()
    e2.record()
    barrier()
    ms_e2e = max_over_ranks(s2.elapsed_time(e2) / args.steps)

    if rank == 0:
        pk = peaks()
        achieved = gemm_flops / gemm_ms / 1e9 if gemm_ms > 0 else 0.0  # TFLOP/s over all GEMM launches in the region
        line = {
            "metric": "frames_per_sec_4step_81f", "value": FRAMES / (DENOISE_STEPS * ms_step * 1e-3), "unit": "frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "int8 (W8A8 GEMM, QK^T) / bf16 (PV, io) / fp32 (accumulate)",
            "data": "synthetic",
            "config": {"workload": f"{shape['name']} DiT denoise step: {layers} blocks (SageSLA top-k {args.topk} + W8A8 + FastNorm)",
                       "L": L, "dim": dim, "heads": heads, "head_dim": d, "ffn": ffn, "text_len": text,
                       "parallelism": f"sp{world}-{sp_mode}" if world > 1 else "single",
                       "l2": "activations per block exceed L2 (>=100 MB tensors, 30+ distinct weight sets)"},
            "e2e": {"value": FRAMES / (DENOISE_STEPS * ms_e2e * 1e-3), "unit": "frames/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": x_host.numel() * 2 + e0_host.numel() * 4 + ctx_host.numel() * 2,
                    "d2h_bytes_per_step": out_host.numel() * 2},
            "gpu_launches": launches,
            "roofline": {"bound": "tensor", "kernel": "gemm_w8a8_kernel (tcgen05 kind::i8)", "achieved": achieved,
                         "peak": pk.get("bf16_tflops_sustained"), "unit": "TFLOP/s",
                         "frac": achieved / pk["bf16_tflops_sustained"] if pk.get("bf16_tflops_sustained") else None,
                         "peak_source": f"{pk['_source']} bf16 sustained (no INT8 peak in MEASURED_PEAKS.json; INT8 nominal is 2x bf16)",
                         "launches": n_gemm, "share_of_step": gemm_share, **ncu_traffic(),
                         "timed_in": "eager pass of the same K steps (CUDA events around each GEMM launch on the launch stream)"},
            "ms_per_step_eager": ms_eager, "launch_mode": graph_note,
            "clocks": clk.summary(),
        }
        if not args.no_cpu_baseline and world == 1:
            r = cpu_reference_sample(shape)
            line["cpu_baseline"] = {"value": FRAMES / (DENOISE_STEPS * r["step_s"]), "unit": "frames/s", "cores": r["cores"],
                                    "kind": "port", "sample": r["sample"]}
        print(json.dumps(line), flush=True)
    if world > 1:
        # NCCL communicators that were captured into a CUDA graph can block in destroy_process_group(); every rank has
        # finished its work here, so release the graph, synchronise, meet at a barrier and leave without the teardown.
        graph = None
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
