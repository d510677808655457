#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_ws -s 4 -c 4 -f -o gpurun_out/prof_gemm python tools/ncu_gemm.py > gpurun_out/ncu_gemm.log 2>&1; echo "ncu_gemm rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_70b.csv python tools/profile_step.py 70b > gpurun_out/ncu_70b.log 2>&1; echo "ncu_70b rc=$?"
ls -la gpurun_out/*.ncu-rep
