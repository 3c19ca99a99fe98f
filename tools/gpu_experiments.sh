#!/bin/bash
# Round-2 kernel experiments prepared (compile-verified) in round 1; run under gpurun on ONE GPU:
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_experiments.sh'
# Each variant is a build-time switch (see the notes at the top of csrc/gemm_w8a8.cu and csrc/sla_attn.cu); the script
# rebuilds only the touched source, runs the parity tests of that kernel, then the microbench lines, and finally restores
# the default build.  Outputs: gpurun_out/exp_<name>_{build,tests}.log, gpurun_out/exp_<name>_mb.jsonl.
mkdir -p gpurun_out
export TDB200_TEST_EXTENDED=1   # include the lazy-rescale attention test (first GPU run pending)
variant() {  # name  source  defines  pytest-args  microbench-filter
  local name=$1 src=$2 defs=$3 tests=$4 filt=$5
  touch turbodiffusion_b200/csrc/$src
  TDB200_NVCC_DEFINES="$defs" python -m turbodiffusion_b200._build > gpurun_out/exp_${name}_build.log 2>&1 || { echo "$name: build failed"; return; }
  timeout 240 python -m pytest $tests -x -q -m gpu > gpurun_out/exp_${name}_tests.log 2>&1
  echo "== $name tests rc=$? $(tail -n 1 gpurun_out/exp_${name}_tests.log)"
  if [ "$filt" = "attn_sweep" ]; then   # fused attention: time vs selected key blocks + the softmax-warp phase trace
    timeout 400 python tools/attn_sweep.py 2>/dev/null | grep '^{' > gpurun_out/exp_${name}_mb.jsonl
    cat gpurun_out/exp_${name}_mb.jsonl
  else
    timeout 400 python tools/microbench.py --filter "$filt" --iters 10 2>/dev/null | grep '^{' > gpurun_out/exp_${name}_mb.jsonl
    python - <<PY
import json
for ln in open("gpurun_out/exp_${name}_mb.jsonl"):
    d = json.loads(ln); print("   ", d["name"], d["ms_median"], "ms", d.get("tflops"), "TFLOP/s")
PY
  fi
}
variant gemm_base    gemm_w8a8.cu ""                         "tests/test_gpu_quant_gemm.py tests/test_gpu_vs_reference_ext.py" "gemm_w8a8/"
variant gemm_cvtmix  gemm_w8a8.cu "-DTDB_GEMM_CVT_MIX=1"     "tests/test_gpu_quant_gemm.py tests/test_gpu_vs_reference_ext.py" "gemm_w8a8/"
variant gemm_ldpipe  gemm_w8a8.cu "-DTDB_GEMM_LD_PIPE=1"     "tests/test_gpu_quant_gemm.py tests/test_gpu_vs_reference_ext.py" "gemm_w8a8/"
variant gemm_ldpipe_cvtmix gemm_w8a8.cu "-DTDB_GEMM_LD_PIPE=1 -DTDB_GEMM_CVT_MIX=1" "tests/test_gpu_quant_gemm.py tests/test_gpu_vs_reference_ext.py" "gemm_w8a8/"
variant attn_base    sla_attn.cu  ""                         "tests/test_gpu_sla.py -k forward"                                "attn_sweep"
variant attn_poly    sla_attn.cu  "-DTDB_ATTN_POLY_EXP2=1"   "tests/test_gpu_sla.py -k forward"                                "attn_sweep"
variant attn_ptmem   sla_attn.cu  "-DTDB_ATTN_P_TMEM=1"      "tests/test_gpu_sla.py -k forward"                                "attn_sweep"
variant attn_ptmem_poly sla_attn.cu "-DTDB_ATTN_P_TMEM=1 -DTDB_ATTN_POLY_EXP2=1" "tests/test_gpu_sla.py -k forward"                       "attn_sweep"
touch turbodiffusion_b200/csrc/gemm_w8a8.cu turbodiffusion_b200/csrc/sla_attn.cu
python -m turbodiffusion_b200._build > gpurun_out/exp_restore_build.log 2>&1 && echo "default build restored"
