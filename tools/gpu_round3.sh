#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() { name=$1; shift; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" | tee -a gpurun_out/summary.txt; tail -3 gpurun_out/$name.log; }
run test_engine python -m pytest tests/test_engine_gpu.py -q -x --no-header -p no:cacheprovider
run bench_8b python bench.py --workload 8b --steps 32 --warmup 4 --no-cpu
run bench_70b python bench.py --workload 70b --steps 24 --warmup 4
run ncu_8b ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_8b.csv python tools/profile_step.py 8b
cat gpurun_out/summary.txt
