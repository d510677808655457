#!/bin/bash
# experiment: unified smem carve-out / smaller attention q tiles, on a 70B-dimension target with 8 layers
mkdir -p gpurun_out
timeout 120 python tools/bench_attn.py > gpurun_out/bench_attn.log 2>&1; echo "bench_attn rc=$?"; cat gpurun_out/bench_attn.log
for v in "base:SSDK_CARVEOUT=-1" "carve:SSDK_CARVEOUT=100" "carve_tq4:SSDK_CARVEOUT=100 SSDK_ATTN_TQ=4" "base_tq4:SSDK_CARVEOUT=-1 SSDK_ATTN_TQ=4"; do
  name=${v%%:*}; envs=${v#*:}
  env $envs timeout 150 python tools/trace_step.py 70b:8 > gpurun_out/trace_$name.log 2>&1; echo "trace $name rc=$?"
  grep -v Warning gpurun_out/trace_$name.log | tail -22
done
