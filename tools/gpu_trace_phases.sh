#!/bin/bash
# Per-kernel timeline of one speculative step on a 70B-dimension target with 8 layers (quick to set up) + the full 1B
# draft: kernel-level increments with the product library, then the phase gaps inside the small kernels with the
# -DSSDK_TRACE_FINE build (python -m ssd_b200.build --trace-variant).  ~1.5 GPU-minutes.
mkdir -p gpurun_out
timeout 150 python tools/trace_step.py 70b:8 > gpurun_out/trace_kernels.log 2>&1; echo "trace rc=$?"
grep -v Warning gpurun_out/trace_kernels.log | grep "step span\|avg=\|layer sample\|^forward"
SSDK_LIB=ssd_b200/_lib/libssdk_trace.so timeout 150 python tools/trace_step.py 70b:8 > gpurun_out/trace_phases.log 2>&1; echo "trace (phase marks) rc=$?"
grep -v Warning gpurun_out/trace_phases.log | grep -A8 "phase gaps"
timeout 200 python tools/bench_attn.py > gpurun_out/bench_attn.log 2>&1; echo "bench_attn rc=$?"; cat gpurun_out/bench_attn.log
