#!/bin/bash
# why does tools/check_draft_stream.py (2-layer 8B-width target, max_model_len 2048) time the streaming draft kernel at
# ~1.04 ms per forward when bench.py / trace_step 70b:8 see 0.73 ms?  phase table of the same configuration
mkdir -p gpurun_out
timeout 40 python tools/trace_step.py 8b:2 --max-len 2048 > gpurun_out/r2f8_trace_8b2_2048.txt 2>&1; echo "rc=$?"
grep "^misc\|per layer\|step span" gpurun_out/r2f8_trace_8b2_2048.txt; grep -A13 "draft_stream_kernel, second forward" gpurun_out/r2f8_trace_8b2_2048.txt | tail -13
