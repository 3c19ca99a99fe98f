#!/bin/bash
# round-2 GPU call: full GPU test suite (attention v1), the SLA tests again on the v2 attention kernel (short timeout: first
# run of a new barrier protocol), hardware probes, GEMM microbench, attention sweep v1/v2, one bench line.
mkdir -p gpurun_out
TDB200_ATTN_IMPL=v1 timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r02_t_all.log 2>&1; echo "tests(v1) rc=$?"; tail -n 25 gpurun_out/r02_t_all.log
TDB200_ATTN_IMPL=v2 timeout 240 python -m pytest tests/test_gpu_sla.py -x -q -m gpu -k "forward" > gpurun_out/r02_t_sla_v2.log 2>&1; echo "sla tests(v2) rc=$?"; tail -n 25 gpurun_out/r02_t_sla_v2.log
timeout 200 python tools/hw_probes.py > gpurun_out/hw_probes.json 2> gpurun_out/hw_probes.err; echo "probes rc=$?"; head -c 6000 gpurun_out/hw_probes.json
timeout 300 python tools/microbench.py --filter gemm_w8a8 --iters 10 --out gpurun_out/r02_mb_gemm.jsonl 2>&1 | grep "^{" | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['name'], d['ms_median'], d.get('tflops'))
"
for impl in v1 v2; do echo "attn_sweep $impl"; TDB200_ATTN_IMPL=$impl timeout 120 python tools/attn_sweep.py 2>/dev/null | grep '^{' | head -6; done
TDB200_ATTN_IMPL=v1 timeout 400 python bench.py --steps 5 --warmup 3 --no-extras > gpurun_out/r02_bench_a.log 2>&1; tail -n 3 gpurun_out/r02_bench_a.log
