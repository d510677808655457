#!/bin/bash
# validate the rewritten norm kernel + 8-warp attention (parity), then compare q-tile sizes on the 8-layer 70B-dimension trace
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_engine_gpu.py -m gpu -x -q 2>&1 | tail -6
for v in "tqdef:A=1" "tq4:SSDK_ATTN_TQ=4" "tq2:SSDK_ATTN_TQ=2"; do
  name=${v%%:*}; envs=${v#*:}
  env $envs timeout 150 python tools/trace_step.py 70b:8 > gpurun_out/trace_$name.log 2>&1; echo "== trace $name rc=$?"
  grep -v Warning gpurun_out/trace_$name.log | grep "step span\|avg=\|target layer\|draft layer\|^forward 2\|^forward 9"
done
SSDK_LIB=ssd_b200/_lib/libssdk_trace.so timeout 150 python tools/trace_step.py 70b:8 > gpurun_out/trace_fine.log 2>&1; echo "== trace fine rc=$?"
grep -v Warning gpurun_out/trace_fine.log | grep -A8 "phase gaps"
