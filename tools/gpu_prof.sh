#!/bin/bash
# ncu captures (1 GPU).  Produces gpurun_out/launches.csv (+ .ncu-rep files for the GEMM and attention kernels).
mkdir -p gpurun_out
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
tail -n 1 gpurun_out/bench.log | cut -c1-300
timeout 700 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 1 --profile > gpurun_out/ncu_launch.log 2>&1; echo "launchlist rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_w8a8 -s 8 -c 2 -f -o gpurun_out/prof_gemm \
   python tools/microbench.py --filter gemm_w8a8/A/ffn --iters 2 > gpurun_out/ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sla_attn -s 1 -c 1 -f -o gpurun_out/prof_attn \
   python bench.py --steps 1 --warmup 1 --layers 2 --profile > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn rc=$?"
ls -la gpurun_out | head -20
