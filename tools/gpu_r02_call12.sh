#!/bin/bash
# round-2 GPU call 12 (1 GPU): whole GPU suite after the INT8-K exchange / split-output GEMM / fused q,k,v changes; step time with and
# without the fused q/k/v GEMM
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -rs > gpurun_out/r02_t_all12.log 2>&1; echo "gpu suite rc=$?"; tail -n 12 gpurun_out/r02_t_all12.log | cut -c1-250
timeout 300 python bench.py --no-extras --steps 5 > gpurun_out/r02_bench_fuseqkv1.log 2>&1; echo "bench fused rc=$?"; grep '^{' gpurun_out/r02_bench_fuseqkv1.log | tail -1 | cut -c1-330
TDB200_FUSE_QKV=0 timeout 300 python bench.py --no-extras --steps 5 > gpurun_out/r02_bench_fuseqkv0.log 2>&1; echo "bench unfused rc=$?"; grep '^{' gpurun_out/r02_bench_fuseqkv0.log | tail -1 | cut -c1-330
