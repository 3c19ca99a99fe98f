#!/bin/bash
# round-2 GPU call 6: whole suite, fused-epilogue rewrite check (microbench + FFN modes), GEMM ncu capture
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -rs > gpurun_out/r02_t_all6.log 2>&1; echo "tests rc=$?"; tail -n 12 gpurun_out/r02_t_all6.log | cut -c1-250
timeout 200 python tools/microbench.py --filter "gelu_quant,gemm_w8a8/A,gemm_w8a8_bias/A,gemm_w8a8_gelu,sla_prep/A/,sla_moments/A/,sla_block_map/A/" --iters 8 --out gpurun_out/r02_mb_call6.jsonl 2>/dev/null | grep '^{' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['name'], d['ms_median'], d.get('tflops'), d.get('gbs'))
"
for mode in split fused; do echo "== bench shape A, FFN activation $mode"; TDB200_FFN_ACT=$mode timeout 300 python bench.py --steps 4 --warmup 3 --no-extras --no-cpu-baseline 2>&1 | grep '^{' | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print(d['ms_per_step'], d['ms_per_step_eager'], d['roofline']['achieved'], d['roofline']['share_of_step'], d['roofline_attention']['achieved'], d['roofline_attention']['share_of_step'], d['gpu_launches'])
"; done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_w8a8 -s 3 -c 2 -f -o gpurun_out/r02_prof_gemm \
   python tools/microbench.py --filter gemm_w8a8/A/ffn_down --iters 2 > gpurun_out/r02_ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
