#!/bin/bash
# Multi-GPU measurements still open after round 1 (run under `gpurun --gpus N`; charged N x box time, keep them short):
#   gpurun --gpus 4 --timeout 600 -- 'bash tools/gpu_scaling.sh 4'
#   gpurun --gpus 8 --timeout 600 -- 'bash tools/gpu_scaling.sh 8'
# N=4: shape A in both exchange modes (only all-gather was measured: 51.1 ms);  N=8: shape A with the uneven-head Ulysses
# split (12 heads -> 2,2,2,2,1,1,1,1; all-gather measured 40.2 ms), plus the 2-GPU parity test of that split.
N=${1:-4}
mkdir -p gpurun_out
run() {  # tag, extra bench args
  timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus $N --steps 3 --warmup 3 $2 > gpurun_out/scale_$1.log 2>&1
  echo "== $1 rc=$? $(tail -n 1 gpurun_out/scale_$1.log | cut -c1-220)"
}
TDB200_TEST_UNEVEN_HEADS=1 timeout 200 python -m pytest tests/test_gpu_dist.py -q -m gpu > gpurun_out/t_dist.log 2>&1; echo "dist tests rc=$? $(tail -n 1 gpurun_out/t_dist.log)"
run A_n${N}_allgather "--sp-mode allgather"
run A_n${N}_ulysses   "--sp-mode ulysses"
