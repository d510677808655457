#!/bin/bash
# round-2 call 1: environment probe for the reference GPU arm + first hardware run of the persistent draft forward
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/r2_smi.txt 2>&1
nproc > gpurun_out/r2_nproc.txt; cat /sys/fs/cgroup/cpu.max >> gpurun_out/r2_nproc.txt 2>&1; python -c "import os; print(len(os.sched_getaffinity(0)))" >> gpurun_out/r2_nproc.txt
timeout 300 python tools/probe_ref_env.py > gpurun_out/r2_probe_ref_env.log 2>&1; echo "probe rc=$?"; tail -6 gpurun_out/r2_probe_ref_env.log
bash tools/gpu_persistent.sh
# reference GPU arm bring-up: tiny, then 8B+1B (each in its own process, bounded)
timeout 600 python baseline/ref_gpu.py --workload tiny --steps 8 --warmup 3 --out gpurun_out/r2_refgpu_tiny.json > gpurun_out/r2_refgpu_tiny.log 2>&1; echo "refgpu tiny rc=$?"; tail -5 gpurun_out/r2_refgpu_tiny.log | cut -c1-600
timeout 900 python baseline/ref_gpu.py --workload 8b --steps 24 --warmup 4 --out gpurun_out/r2_refgpu_8b.json > gpurun_out/r2_refgpu_8b.log 2>&1; echo "refgpu 8b rc=$?"; tail -3 gpurun_out/r2_refgpu_8b.log | cut -c1-600
