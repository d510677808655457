#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() { name=$1; shift; timeout 300 "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" | tee -a gpurun_out/summary.txt; tail -3 gpurun_out/$name.log | cut -c1-600; }
run test_ops python -m pytest tests/test_ops_gpu.py -q -x --no-header -p no:cacheprovider -k "paged_attention or rms_norm or rope"
run test_engine python -m pytest tests/test_engine_gpu.py -q -x --no-header -p no:cacheprovider
run trace_8b python tools/trace_step.py 8b
tail -22 gpurun_out/trace_8b.log
run trace_70b python tools/trace_step.py 70b
tail -22 gpurun_out/trace_70b.log
cat gpurun_out/summary.txt
