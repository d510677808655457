#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/trace_step.py 8b > gpurun_out/trace_8b.log 2>&1; echo "trace_8b rc=$?"; tail -40 gpurun_out/trace_8b.log
timeout 300 python tools/trace_step.py 70b > gpurun_out/trace_70b.log 2>&1; echo "trace_70b rc=$?"; tail -30 gpurun_out/trace_70b.log
