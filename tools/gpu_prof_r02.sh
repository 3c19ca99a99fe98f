#!/bin/bash
# round-2 ncu captures (1 GPU): launch list of one profiled step, `--set full` of the W8A8 GEMM (ffn-down A) and of the fused
# attention kernel.  The .ncu-rep files stay in gpurun_out/; tools/ncu_summary.py turns them into profiles/r02_*.txt here.
mkdir -p gpurun_out
timeout 700 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches.csv \
   python bench.py --steps 1 --warmup 1 --layers 3 --profile > gpurun_out/r02_ncu_launch.log 2>&1; echo "launchlist rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_w8a8 -s 3 -c 2 -f -o gpurun_out/r02_prof_gemm \
   python tools/microbench.py --filter gemm_w8a8/A/ffn_down --iters 2 > gpurun_out/r02_ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sla_attn -s 1 -c 1 -f -o gpurun_out/r02_prof_attn \
   python bench.py --steps 1 --warmup 1 --layers 2 --profile > gpurun_out/r02_ncu_attn.log 2>&1; echo "ncu attn rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:pool_quant|sla_moments|row_norm|ln_modulate|block_map" -s 12 -c 10 -f -o gpurun_out/r02_prof_prologue \
   python bench.py --steps 1 --warmup 1 --layers 2 --profile > gpurun_out/r02_ncu_prologue.log 2>&1; echo "ncu prologue rc=$?"
ls -la gpurun_out | grep r02_prof
