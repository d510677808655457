#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() { name=$1; shift; timeout 400 "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" | tee -a gpurun_out/summary.txt; tail -3 gpurun_out/$name.log | cut -c1-400; }
run test_all python -m pytest tests -q -x --no-header -p no:cacheprovider -m gpu
run trace_8b python tools/trace_step.py 8b
grep -E "^gemm|^norm|^misc|^attn|^rope|layer sample|span" gpurun_out/trace_8b.log | cut -c1-420
run trace_70b python tools/trace_step.py 70b
grep -E "^gemm|^norm|^misc|^attn|^rope|layer sample|span" gpurun_out/trace_70b.log | cut -c1-420
bash tools/gpu_ncu.sh
cat gpurun_out/summary.txt
