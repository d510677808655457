#!/bin/bash
# round 2, closing call: prefill timings (chunk 64 vs 256; 16 x 128-token prompts packed) and the default bench line with the
# reference GPU arm from the final build
mkdir -p gpurun_out
timeout 400 python tools/bench_prefill.py > gpurun_out/r2f6_prefill.txt 2> gpurun_out/r2f6_prefill.err; echo "prefill rc=$?"; cat gpurun_out/r2f6_prefill.txt
( time timeout 900 python bench.py > gpurun_out/r2f6_bench_70b_tp1.json 2> gpurun_out/r2f6_bench_70b_tp1.err ) 2>&1 | tail -3
echo "bench rc=$?"; cut -c1-330 gpurun_out/r2f6_bench_70b_tp1.json; grep -o '"reference_gpu": {[^}]*}' gpurun_out/r2f6_bench_70b_tp1.json | cut -c1-600
