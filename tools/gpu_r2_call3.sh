#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -v "Warning\|warn" | tail -40 ) > gpurun_out/r2_pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2_pytest_gpu.txt
( time timeout 900 python bench.py > gpurun_out/r2_bench_default.log 2>gpurun_out/r2_bench_default.err ); echo "bench default rc=$?"; tail -c 1500 gpurun_out/r2_bench_default.log
( time timeout 600 python bench.py --workload 8b --steps 32 --warmup 4 --no-cpu > gpurun_out/r2_bench_8b.log 2>gpurun_out/r2_bench_8b.err ); echo "bench 8b rc=$?"; tail -c 1200 gpurun_out/r2_bench_8b.log
bash tools/gpu_ncu_families.sh
