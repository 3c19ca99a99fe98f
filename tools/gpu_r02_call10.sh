#!/bin/bash
# round-2 GPU call 10 (1 GPU): SLA tests after the moments-kernel rewrite, moments microbench, the full default bench line (wall clock)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sla.py tests/test_gpu_fullsize.py tests/test_gpu_fullsize_parity.py tests/test_gpu_block.py -q -m gpu > gpurun_out/r02_t_sla10.log 2>&1; echo "sla tests rc=$?"; tail -n 5 gpurun_out/r02_t_sla10.log | cut -c1-250
timeout 200 python tools/microbench.py --filter "sla_moments/A/,sla_prep/A/,sla_moments/A64,sla_prep/A64,sla_moments/B/" --iters 8 --out gpurun_out/r02_mb_call10.jsonl 2>/dev/null | grep '^{' | grep -E "moments|prep" | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['name'], d['ms_median'], d.get('tflops'), d.get('gbs'), d.get('frac_hbm_peak'))
"
SECONDS=0; timeout 900 python bench.py > gpurun_out/r02_bench_default.log 2> gpurun_out/r02_bench_default.err; echo "bench rc=$? wall=${SECONDS}s"; grep '^{' gpurun_out/r02_bench_default.log | tail -1 | cut -c1-5000; tail -3 gpurun_out/r02_bench_default.err | cut -c1-300
