#!/bin/bash
# round-2 GPU call 5: whole GPU suite (no -x: list every failure), attention timeline to a file, ncu captures
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -rs > gpurun_out/r02_t_all5.log 2>&1; echo "tests rc=$?"; tail -n 40 gpurun_out/r02_t_all5.log | cut -c1-250
timeout 120 python tools/attn_sweep.py 2>/dev/null | grep '^{' > gpurun_out/r02_attn_sweep.jsonl; cat gpurun_out/r02_attn_sweep.jsonl | cut -c1-900
bash tools/gpu_prof_r02.sh
