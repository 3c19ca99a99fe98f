#!/bin/bash
# round-2 GPU call 5: whole GPU suite (no -x: list every failure), attention timeline to a file, ncu captures
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -rs > gpurun_out/r02_t_all5.log 2>&1; echo "tests rc=$?"; tail -n 40 gpurun_out/r02_t_all5.log | cut -c1-250
timeout 120 python tools/attn_sweep.py 2>/dev/null | grep '^{' > gpurun_out/r02_attn_sweep.jsonl; cat gpurun_out/r02_attn_sweep.jsonl | cut -c1-900
bash tools/gpu_prof_r02.sh
timeout 200 python tools/microbench.py --filter "gelu_quant,gemm_w8a8/A,gemm_w8a8_bias/A,gemm_w8a8_gelu" --iters 8 --out gpurun_out/r02_mb_call5.jsonl 2>/dev/null | grep '^{' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['name'], d['ms_median'], d.get('tflops'), d.get('gbs'))
"
for mode in split fused; do echo "== bench shape A, FFN activation $mode"; TDB200_FFN_ACT=$mode timeout 300 python bench.py --steps 4 --warmup 3 --no-extras --no-cpu-baseline 2>&1 | grep '^{' | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print(d['ms_per_step'], d['ms_per_step_eager'], d['roofline']['achieved'], d['roofline']['share_of_step'], d['roofline_attention']['achieved'], d['roofline_attention']['share_of_step'], d['gpu_launches'])
"; done
