#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/check_draft_stream.py > gpurun_out/r2_stream_check.log 2>&1; echo "stream check rc=$?"; tail -1 gpurun_out/r2_stream_check.log | cut -c1-600
( time timeout 1200 python bench.py > gpurun_out/r2_bench_70b_tp1.json 2> gpurun_out/r2_bench_70b_tp1.err ) 2>&1 | tail -3; echo "bench default rc=$?"; cut -c1-330 gpurun_out/r2_bench_70b_tp1.json
( time timeout 900 python bench.py --workload 8b --steps 48 --warmup 6 > gpurun_out/r2_bench_8b_tp1.json 2> gpurun_out/r2_bench_8b_tp1.err ) 2>&1 | tail -3; cut -c1-330 gpurun_out/r2_bench_8b_tp1.json
( time timeout 900 python bench.py --workload qwen32b --temp 0.7 --lm-scale 10 --steps 48 --warmup 6 --no-cpu > gpurun_out/r2_bench_qwen32b_tp1_t07.json 2> gpurun_out/r2_bench_qwen32b_tp1_t07.err ) 2>&1 | tail -3; cut -c1-330 gpurun_out/r2_bench_qwen32b_tp1_t07.json
timeout 300 python tools/trace_step.py 70b > gpurun_out/r2_timeline_70b_tp1.txt 2>&1; echo "trace 70b rc=$?"; grep -v Warning gpurun_out/r2_timeline_70b_tp1.txt | grep "step span\|avg=\|^forward" | tail -12
timeout 300 python tools/trace_step.py 8b > gpurun_out/r2_timeline_8b_tp1.txt 2>&1; echo "trace 8b rc=$?"; grep -v Warning gpurun_out/r2_timeline_8b_tp1.txt | grep "step span\|avg=" | head -8
