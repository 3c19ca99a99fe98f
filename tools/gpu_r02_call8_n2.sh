#!/bin/bash
# round-2 GPU call 8 (2 GPUs): sequence-parallel parity tests and the bench line at N=2 (with extra_configs: shape B at N=2)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist.py -q -m gpu -rs > gpurun_out/r02_t_dist.log 2>&1; echo "dist tests rc=$?"; tail -n 8 gpurun_out/r02_t_dist.log | cut -c1-250
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_n2.log 2>&1; echo "bench n2 rc=$?"; grep '^{' gpurun_out/r02_bench_n2.log | tail -1 | cut -c1-1500
