#!/bin/bash
# round-2 GPU call 3: GEMM variant matrix, attention v1 timeline (2 CTAs/SM vs 1 CTA/SM), remaining tests, row-norm microbench
mkdir -p gpurun_out
bash tools/gpu_gemm_variants.sh
echo "== attention v1, two CTAs per SM"; TDB200_ATTN_IMPL=v1 timeout 120 python tools/attn_sweep.py 2>/dev/null | grep '^{'
echo "== attention v1, ONE CTA per SM"; TDB200_ATTN_IMPL=v1 TDB200_ATTN_ONE_CTA_PER_SM=1 timeout 120 python tools/attn_sweep.py 2>/dev/null | grep '^{'
TDB200_ATTN_IMPL=v1 timeout 900 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_quant_gemm.py > gpurun_out/r02_t_rest.log 2>&1; echo "tests(rest, v1) rc=$?"; tail -n 30 gpurun_out/r02_t_rest.log
timeout 200 python tools/microbench.py --filter "norm,quant_int8,ln_modulate,gate_residual" --iters 8 --out gpurun_out/r02_mb_norm.jsonl 2>/dev/null | grep '^{' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['name'], d['ms_median'], d.get('gbs'), d.get('frac_hbm_peak'))
"
for mode in split fused; do echo "== bench shape A, FFN activation $mode"; TDB200_ATTN_IMPL=v1 TDB200_FFN_ACT=$mode timeout 300 python bench.py --steps 4 --warmup 3 --no-extras --no-cpu-baseline 2>&1 | grep '^{' | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print(d['ms_per_step'], d['ms_per_step_eager'], d['roofline']['achieved'], d['roofline']['share_of_step'], d['roofline_attention']['achieved'], d['roofline_attention']['share_of_step'], d['gpu_launches'])
"; done
