#!/bin/bash
# One `ncu --set full` capture per kernel family of the hot path (north_star: "each kernel ships with an ncu capture").
# ncu serialises kernels and flushes caches between replays: durations are cold-cache; they document traffic and
# pipeline behaviour, the in-graph timeline (tools/trace_step.py) documents the critical path.
# gpurun brings back at most 64 MiB: the raw metric pages are exported to CSV on the box and only the attention and GEMM
# reports travel as .ncu-rep.
mkdir -p gpurun_out/ncu
NCU="ncu --set full --clock-control none -f"
cap() {  # name regex skip count mode keep_rep [env]
  timeout 600 env $7 $NCU -k regex:$2 -s $3 -c $4 -o gpurun_out/ncu/r02_$1 python tools/ncu_workload.py $5 > gpurun_out/ncu/r02_$1.log 2>&1
  echo "ncu $1 rc=$?"
  ncu -i gpurun_out/ncu/r02_$1.ncu-rep --page raw --csv > gpurun_out/ncu/r02_$1.raw.csv 2>/dev/null
  [ "$6" = "keep" ] || rm -f gpurun_out/ncu/r02_$1.ncu-rep
}
cap attn       'paged_attn_kernel|attn_combine_kernel' 0 24 attn keep
cap gemm       'gemm_ws_kernel|splitk_reduce'          8 8  gemm keep
# engine workload: skip the prefill launches, capture from inside the spec steps
cap norm       'add_rmsnorm_kernel'  60 6 engine drop
cap rope       'rope_store_kernel'   30 4 engine drop
cap sample     'sample_kernel'       1  3 engine drop SSDK_DRAFT_STREAM=0
cap verify     'verify_kernel'       0  2 engine drop
cap verify_t07 'verify_kernel'       0  2 engine drop NCU_TEMP=0.7
cap sample_t07 'sample_kernel'       1  3 engine drop "NCU_TEMP=0.7 SSDK_DRAFT_STREAM=0"
cap draftattn  'paged_attn_kernel'   30 4 engine drop SSDK_DRAFT_STREAM=0
cap stream     'draft_stream_kernel' 1  1 engine drop
du -sh gpurun_out/ncu; ls -la gpurun_out/ncu/
