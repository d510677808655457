#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_persistent.sh
timeout 300 python bench.py --workload 8b --steps 16 --warmup 4 --no-cpu > gpurun_out/bench_8b_regular.log 2>&1; echo "bench 8b regular rc=$?"
grep '^{' gpurun_out/bench_8b_regular.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['accept_len'])"
timeout 600 python baseline/ref_gpu.py --workload tiny --steps 8 --warmup 3 --out gpurun_out/r2_refgpu_tiny.json > gpurun_out/r2_refgpu_tiny.log 2>&1; echo "refgpu tiny rc=$?"; tail -5 gpurun_out/r2_refgpu_tiny.log | cut -c1-600
timeout 900 python baseline/ref_gpu.py --workload 8b --steps 24 --warmup 4 --out gpurun_out/r2_refgpu_8b.json > gpurun_out/r2_refgpu_8b.log 2>&1; echo "refgpu 8b rc=$?"; tail -3 gpurun_out/r2_refgpu_8b.log | cut -c1-600
