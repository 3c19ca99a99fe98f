#!/bin/bash
# round-2 GPU call 16 (1 GPU): whole GPU suite (RoPE table, fused cross-attention k/v, seeded RNG), rmsnorm_rope microbench with / without
# the table, the attention phase trace with the S(j+1) / P.V(j-1) waits separated, step time with / without the table
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -rs > gpurun_out/r02_t_all16.log 2>&1; echo "gpu suite rc=$?"; tail -n 6 gpurun_out/r02_t_all16.log | cut -c1-250
for t in 1 0; do TDB200_ROPE_TABLE=$t timeout 120 python tools/microbench.py --filter "rmsnorm_rope/" --iters 10 --out gpurun_out/r02_mb_rope_table$t.jsonl 2>/dev/null | grep '^{' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('table=$t', d['name'], d['ms_median'], d.get('frac_hbm_peak'))
"; done
timeout 200 python tools/attn_sweep.py > gpurun_out/r02_attn_sweep2.jsonl 2>gpurun_out/attn_sweep2.err; echo "sweep rc=$?"; tail -n 2 gpurun_out/r02_attn_sweep2.jsonl | cut -c1-900
timeout 300 python bench.py --no-extras --steps 5 > gpurun_out/r02_bench_ropetable1.log 2>&1; echo "bench table rc=$?"; grep '^{' gpurun_out/r02_bench_ropetable1.log | tail -1 | cut -c1-330
TDB200_ROPE_TABLE=0 timeout 300 python bench.py --no-extras --steps 5 > gpurun_out/r02_bench_ropetable0.log 2>&1; echo "bench sincos rc=$?"; grep '^{' gpurun_out/r02_bench_ropetable0.log | tail -1 | cut -c1-330
