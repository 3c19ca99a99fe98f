#!/bin/bash
# round-2 GPU call 9 (1 GPU): the reference WanModel step on these operators (f1c) and the full default bench line (with extras)
mkdir -p gpurun_out
timeout 900 python tools/ref_model_step.py > gpurun_out/r02_ref_model_step.json 2> gpurun_out/r02_ref_model_step.err; echo "ref model rc=$?"; cat gpurun_out/r02_ref_model_step.json; tail -3 gpurun_out/r02_ref_model_step.err | cut -c1-300
/usr/bin/time -v timeout 900 python bench.py > gpurun_out/r02_bench_default.log 2> gpurun_out/r02_bench_default.err; echo "bench rc=$?"; grep '^{' gpurun_out/r02_bench_default.log | tail -1 | cut -c1-4000; grep -E "Elapsed|Maximum resident" gpurun_out/r02_bench_default.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.log 2>&1; tail -1 gpurun_out/r02_bench_reference.log | cut -c1-1200
