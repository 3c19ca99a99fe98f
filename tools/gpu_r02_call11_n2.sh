#!/bin/bash
# round-2 GPU call 11 (2 GPUs): split-key-prep tests, sequence-parallel parity tests (both modes, uneven heads), bench at N=2:
# INT8-K all-gather (default, with extra_configs) vs the 16-bit K exchange vs the head<->sequence all-to-all
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_sla.py -q -m gpu -k "split_key or sequence_major" > gpurun_out/r02_t_kseq.log 2>&1; echo "kseq tests rc=$?"; tail -n 4 gpurun_out/r02_t_kseq.log | cut -c1-250
TDB200_TEST_UNEVEN_HEADS=1 timeout 600 python -m pytest tests/test_gpu_dist.py -q -m gpu -rs > gpurun_out/r02_t_dist.log 2>&1; echo "dist tests rc=$?"; tail -n 8 gpurun_out/r02_t_dist.log | cut -c1-250
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_n2.log 2>&1; echo "bench n2 rc=$?"; grep '^{' gpurun_out/r02_bench_n2.log | tail -1 | cut -c1-1200; tail -3 gpurun_out/r02_bench_n2.log | grep -v '^{' | cut -c1-300
TDB200_SP_INT8_K=0 timeout 600 $TR --master-port 29542 bench.py --gpus 2 --steps 5 --warmup 3 --no-extras > gpurun_out/r02_bench_n2_bf16k.log 2>&1; echo "bench n2 bf16-K rc=$?"; grep '^{' gpurun_out/r02_bench_n2_bf16k.log | tail -1 | cut -c1-400
timeout 600 $TR --master-port 29543 bench.py --gpus 2 --steps 5 --warmup 3 --no-extras --sp-mode ulysses > gpurun_out/r02_bench_n2_ulysses.log 2>&1; echo "bench n2 ulysses rc=$?"; grep '^{' gpurun_out/r02_bench_n2_ulysses.log | tail -1 | cut -c1-400
