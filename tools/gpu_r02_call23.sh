#!/bin/bash
# round-2 GPU call 23 (1 GPU): whole GPU suite after the launch-count changes (moments projection kernel, one zero fill, per-step
# modulation add), then the default bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -rs > gpurun_out/r02_t_all23.log 2>&1; echo "gpu suite rc=$?"; tail -n 6 gpurun_out/r02_t_all23.log | cut -c1-250
SECONDS=0; timeout 900 python bench.py > gpurun_out/r02_bench_default2.log 2> gpurun_out/r02_bench_default2.err; echo "bench rc=$? wall=${SECONDS}s"; grep '^{' gpurun_out/r02_bench_default2.log | tail -1 | cut -c1-700
