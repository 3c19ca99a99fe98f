#!/bin/bash
# round-2 GPU call 7: suite after the conversion-unit changes (magic-add int8 conversion, packed roundings, FRND-free RoPE),
# prologue/quant microbench, bench in both FFN modes
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -rs > gpurun_out/r02_t_all7.log 2>&1; echo "tests rc=$?"; tail -n 12 gpurun_out/r02_t_all7.log | cut -c1-250
timeout 300 python tools/microbench.py --filter "norm,quant_int8,ln_modulate,gate_residual,gemm_w8a8_gelu,sla_prep/A/,sla_moments/A/" --iters 8 --out gpurun_out/r02_mb_call7.jsonl 2>/dev/null | grep '^{' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['name'], d['ms_median'], d.get('tflops'), d.get('gbs'), d.get('frac_hbm_peak'))
"
for mode in fused split; do echo "== bench shape A, FFN activation $mode"; TDB200_FFN_ACT=$mode timeout 300 python bench.py --steps 4 --warmup 3 --no-extras --no-cpu-baseline 2>&1 | grep '^{' | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print(d['ms_per_step'], d['ms_per_step_eager'], d['roofline']['achieved'], d['roofline']['share_of_step'], d['roofline_attention']['achieved'], d['roofline_attention']['share_of_step'], d['gpu_launches'])
"; done
