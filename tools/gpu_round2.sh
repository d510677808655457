#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() { name=$1; shift; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" | tee -a gpurun_out/summary.txt; tail -4 gpurun_out/$name.log; }
run test_engine python -m pytest tests/test_engine_gpu.py -q -x --no-header -p no:cacheprovider
run smoke python __graft_entry__.py smoke
run gemm_bench python tools/bench_gemm.py
run bench_tiny python bench.py --workload tiny --steps 8 --warmup 3 --no-cpu
run bench_8b python bench.py --workload 8b --steps 32 --warmup 4 --no-cpu
run bench_8b_nopdl python bench.py --workload 8b --steps 32 --warmup 4 --no-cpu --no-pdl
cat gpurun_out/summary.txt
