#!/bin/bash
# Sequence-parallel measurements on N GPUs of one box (run under `gpurun --gpus N`; charged N x box time, keep it short):
#   gpurun --gpus 8 --timeout 900 -- 'bash tools/gpu_scaling.sh 8'
# Runs the 2-GPU parity tests when N == 2, then shape A with the mode `auto` picks (with extra_configs: shape B at this N) and with each
# exchange mode forced.  Outputs: gpurun_out/scale_<tag>.log (last line = the bench JSON).
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() {  # tag, extra bench args
  timeout 600 $TR --master-port $((29540 + RANDOM % 400)) bench.py --gpus $N --steps 10 --warmup 3 $2 > gpurun_out/scale_$1.log 2>&1
  echo "== $1 rc=$? $(grep '^{' gpurun_out/scale_$1.log | tail -n 1 | cut -c1-240)"
}
if [ "$N" = "2" ]; then
  timeout 600 python -m pytest tests/test_gpu_dist.py -q -m gpu > gpurun_out/t_dist.log 2>&1; echo "dist tests rc=$? $(tail -n 1 gpurun_out/t_dist.log)"
fi
run A_n${N}_auto ""
run A_n${N}_allgather "--sp-mode allgather --no-extras"
run A_n${N}_ulysses   "--sp-mode ulysses --no-extras"
