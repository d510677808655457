#!/bin/bash
# experiment: GEMM pre-wait prefetch off / pre-wait page touching / fused rope prologue
mkdir -p gpurun_out
for v in "base:SSDK_KNOBS=0" "noprefetch:SSDK_KNOBS=1" "pretouch:SSDK_KNOBS=2" "fuse:SSDK_FUSE_ROPE=1" "fuse_tq4:SSDK_FUSE_ROPE=1 SSDK_ATTN_TQ=4" "fuse_tq4_touch:SSDK_FUSE_ROPE=1 SSDK_ATTN_TQ=4 SSDK_KNOBS=2"; do
  name=${v%%:*}; envs=${v#*:}
  env SSDK_CARVEOUT=-1 $envs timeout 150 python tools/trace_step.py 70b:8 > gpurun_out/trace_$name.log 2>&1; echo "== trace $name rc=$?"
  grep -v Warning gpurun_out/trace_$name.log | grep "step span\|target layer\|draft layer\|^forward 2\|^forward 99\|phase gaps\|^  " 
done
SSDK_FUSE_ROPE=1 timeout 300 python -m pytest tests/test_engine_gpu.py tests/test_llm_gpu.py -m gpu -x -q 2>&1 | tail -5
SSDK_FUSE_ROPE=1 SSDK_ATTN_TQ=4 timeout 300 python -m pytest tests/test_engine_gpu.py -m gpu -x -q 2>&1 | tail -3
