#!/bin/bash
# round-2 GPU call 4: full GPU suite on the new defaults, attention timeline after the epilogue rewrite, FFN activation modes,
# prologue microbench, reference-vs-ours (reference Triton SLA / FastNorm + reference CUDA GEMM) on the same B200.
mkdir -p gpurun_out
timeout 1000 python -m pytest tests -x -q -m gpu > gpurun_out/r02_t_all4.log 2>&1; echo "tests rc=$?"; tail -n 30 gpurun_out/r02_t_all4.log
echo "== attention v1 (new epilogue)"; timeout 120 python tools/attn_sweep.py 2>/dev/null | grep '^{'
timeout 200 python tools/microbench.py --filter "norm,quant_int8,ln_modulate,gate_residual" --iters 8 --out gpurun_out/r02_mb_prologue.jsonl 2>/dev/null | grep '^{' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['name'], d['ms_median'], d.get('gbs'), d.get('frac_hbm_peak'))
"
for mode in split fused; do echo "== bench shape A, FFN activation $mode"; TDB200_FFN_ACT=$mode timeout 300 python bench.py --steps 4 --warmup 3 --no-extras --no-cpu-baseline 2>&1 | grep '^{' | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print(d['ms_per_step'], d['ms_per_step_eager'], d['roofline']['achieved'], d['roofline']['share_of_step'], d['roofline_attention']['achieved'], d['roofline_attention']['share_of_step'], d['gpu_launches'])
"; done
timeout 400 python tools/ref_vs_ours.py > gpurun_out/r02_ref_vs_ours.log 2>&1; echo "ref_vs_ours rc=$?"; grep '^{' gpurun_out/r02_ref_vs_ours.log | cut -c1-330
