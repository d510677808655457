#!/bin/bash
mkdir -p gpurun_out
SSDK_LIB=$PWD/ssd_b200/_lib/libssdk_trace.so timeout 300 python tools/trace_step.py 70b:8 > gpurun_out/r2_trace_stream_probe.log 2>&1; echo "trace probe rc=$?"; grep -v Warning gpurun_out/r2_trace_stream_probe.log | grep -A17 "draft_stream_kernel, second" | head -20
timeout 300 python tools/trace_step.py 70b:8 > gpurun_out/r2_trace_stream4.log 2>&1; echo "trace rc=$?"; grep -v Warning gpurun_out/r2_trace_stream4.log | grep -A16 "draft_stream_kernel, second" | head -18; grep "avg=" gpurun_out/r2_trace_stream4.log | head -2
