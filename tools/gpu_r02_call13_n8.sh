#!/bin/bash
# round-2 GPU call 13 (8 GPUs): the driver's scaling command at N=8 (shape A, mode auto = INT8-K all-gather for 12 heads; extra_configs
# carries shape B at N=8 = head<->sequence all-to-all), then shape A with the fused q/k/v GEMM under the hook, and with the uneven-head
# all-to-all (12 heads -> 2,2,2,2,1,1,1,1)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29551 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_bench_n8.log 2>&1; echo "bench n8 rc=$?"; grep '^{' gpurun_out/r02_bench_n8.log | tail -1 | cut -c1-700; grep -v '^{' gpurun_out/r02_bench_n8.log | tail -3 | cut -c1-300
TDB200_FUSE_QKV=1 timeout 600 $TR --master-port 29552 bench.py --gpus 8 --steps 10 --warmup 3 --no-extras > gpurun_out/r02_bench_n8_fuseqkv.log 2>&1; echo "bench n8 fused-qkv rc=$?"; grep '^{' gpurun_out/r02_bench_n8_fuseqkv.log | tail -1 | cut -c1-330
timeout 600 $TR --master-port 29553 bench.py --gpus 8 --steps 10 --warmup 3 --no-extras --sp-mode ulysses > gpurun_out/r02_bench_n8_ulysses.log 2>&1; echo "bench n8 uneven ulysses rc=$?"; grep '^{' gpurun_out/r02_bench_n8_ulysses.log | tail -1 | cut -c1-330
