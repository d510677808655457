#!/bin/bash
# One `ncu --set full` capture per kernel family of the hot path (north_star: "each kernel ships with an ncu capture").
# ncu serialises kernels and flushes caches between replays: durations are cold-cache; they document traffic and
# pipeline behaviour, the in-graph timeline (tools/trace_step.py) documents the critical path.
mkdir -p gpurun_out/ncu
NCU="ncu --set full --clock-control none --import-source on -f"
cap() {  # name regex skip count mode [env]
  timeout 600 env $6 $NCU -k regex:$2 -s $3 -c $4 -o gpurun_out/ncu/r02_$1 python tools/ncu_workload.py $5 > gpurun_out/ncu/r02_$1.log 2>&1
  echo "ncu $1 rc=$?"
}
cap attn       'paged_attn_kernel|attn_combine_kernel' 0 18 attn
cap gemm       'gemm_ws_kernel|splitk_reduce'          8 8  gemm
# engine workload: skip the prefill launches, capture from inside the spec steps
cap norm       'add_rmsnorm_kernel'  60 8 engine
cap rope       'rope_store_kernel'   30 6 engine
cap sample     'sample_kernel'       1  4 engine
cap verify     'verify_kernel'       0  3 engine
cap verify_t07 'verify_kernel'       0  3 engine NCU_TEMP=0.7
cap sample_t07 'sample_kernel'       1  4 engine NCU_TEMP=0.7
cap draftattn  'paged_attn_kernel'   30 6 engine
ls -la gpurun_out/ncu/
