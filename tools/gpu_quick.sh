#!/bin/bash
# quick iteration loop: attention/engine parity + 8B bench (+ optional ncu launch list with NCU=1)
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() { name=$1; shift; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" | tee -a gpurun_out/summary.txt; tail -3 gpurun_out/$name.log; }
run test_attn python -m pytest tests/test_ops_gpu.py -q -x --no-header -p no:cacheprovider -k "paged_attention or rms_norm or rope or linear"
run test_engine python -m pytest tests/test_engine_gpu.py -q -x --no-header -p no:cacheprovider
run bench_8b python bench.py --workload 8b --steps 32 --warmup 4 --no-cpu
if [ -n "$NCU" ]; then
run ncu_8b ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_8b.csv python tools/profile_step.py 8b
fi
if [ -n "$B70" ]; then
run bench_70b python bench.py --workload 70b --steps 24 --warmup 4 --no-cpu
fi
cat gpurun_out/summary.txt
