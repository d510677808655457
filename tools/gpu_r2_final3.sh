#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python bench.py > gpurun_out/r2_bench_70b_tp1.json 2> gpurun_out/r2_bench_70b_tp1.err ) 2>&1 | tail -3; cut -c1-330 gpurun_out/r2_bench_70b_tp1.json
( time timeout 600 python bench.py --workload 8b --steps 48 --warmup 6 > gpurun_out/r2_bench_8b_tp1.json 2> gpurun_out/r2_bench_8b_tp1.err ) 2>&1 | tail -3; cut -c1-330 gpurun_out/r2_bench_8b_tp1.json
