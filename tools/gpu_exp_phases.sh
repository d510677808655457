#!/bin/bash
# experiment: where do the small kernels spend their time (phase marks) + stand-alone attention timings
mkdir -p gpurun_out
SSDK_CARVEOUT=-1 timeout 150 python tools/trace_step.py 70b:8 > gpurun_out/trace_phases.log 2>&1; echo "trace rc=$?"
grep -v Warning gpurun_out/trace_phases.log | tail -40
timeout 200 python tools/bench_attn.py > gpurun_out/bench_attn.log 2>&1; echo "bench_attn rc=$?"; cat gpurun_out/bench_attn.log
