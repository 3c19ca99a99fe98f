#!/bin/bash
# round-2 GPU call 17 (1 GPU): attention after dropping the redundant P.V-barrier poll: SLA / block / full-size parity tests, sweep + trace,
# kernel and module microbench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sla.py tests/test_gpu_fullsize.py tests/test_gpu_fullsize_parity.py tests/test_gpu_block.py tests/test_gpu_reference_model.py -q -m gpu > gpurun_out/r02_t_sla17.log 2>&1; echo "sla tests rc=$?"; tail -n 4 gpurun_out/r02_t_sla17.log | cut -c1-250
timeout 200 python tools/attn_sweep.py > gpurun_out/r02_attn_sweep3.jsonl 2>gpurun_out/attn_sweep3.err; echo "sweep rc=$?"; cat gpurun_out/r02_attn_sweep3.jsonl | cut -c1-700
timeout 200 python tools/microbench.py --filter "sla_attn/A/,sla_module/A/,sla_attn/B/" --iters 8 --out gpurun_out/r02_mb_call17.jsonl 2>/dev/null | grep '^{' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['name'], d['ms_median'], d.get('tflops'))
"
