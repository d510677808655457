#!/bin/bash
mkdir -p gpurun_out
SSDK_LIB=$PWD/ssd_b200/_lib/libssdk_trace.so timeout 300 python tools/trace_step.py 70b:8 > gpurun_out/r2_trace_stream_probe.log 2>&1; echo "trace probe rc=$?"; grep -v Warning gpurun_out/r2_trace_stream_probe.log | grep -A17 "draft_stream_kernel, second" | head -20
timeout 300 python tools/trace_step.py 70b:8 > gpurun_out/r2_trace_stream6.log 2>&1; echo "trace rc=$?"; grep "avg=" gpurun_out/r2_trace_stream6.log | head -2
( time timeout 1800 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "Warning\|warn" | grep "passed\|failed\|FAILED\|Error\|\[KL\]\|\[true-width\|\[golden trace\|mean_accept_len" | tail -40 ) > gpurun_out/r2_pytest_gpu.txt 2>&1; echo "pytest done"; cat gpurun_out/r2_pytest_gpu.txt | cut -c1-400
bash tools/gpu_ncu_families.sh 2>&1 | tail -25
