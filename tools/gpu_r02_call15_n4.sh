#!/bin/bash
# round-2 GPU call 15 (4 GPUs): shape A at N=4 in both exchange modes (auto = all-gather there; 12 heads split evenly for the all-to-all)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29561 bench.py --gpus 4 --steps 10 --warmup 3 --no-extras > gpurun_out/r02_bench_n4.log 2>&1; echo "bench n4 rc=$?"; grep '^{' gpurun_out/r02_bench_n4.log | tail -1 | cut -c1-330
timeout 600 $TR --master-port 29562 bench.py --gpus 4 --steps 10 --warmup 3 --no-extras --sp-mode ulysses > gpurun_out/r02_bench_n4_ulysses.log 2>&1; echo "bench n4 ulysses rc=$?"; grep '^{' gpurun_out/r02_bench_n4_ulysses.log | tail -1 | cut -c1-330
