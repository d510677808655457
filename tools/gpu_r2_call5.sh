#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/trace_step.py 70b:8 > gpurun_out/r2_trace_stream2.log 2>&1; echo "trace rc=$?"; grep -v Warning gpurun_out/r2_trace_stream2.log | grep -A20 "draft_stream_kernel, second" | head -24; grep "avg=" gpurun_out/r2_trace_stream2.log | head -4
timeout 900 python tools/check_draft_stream.py > gpurun_out/r2_stream_check.log 2>&1; echo "stream check rc=$?"; tail -1 gpurun_out/r2_stream_check.log | cut -c1-900
timeout 400 python bench.py --workload 8b --steps 24 --warmup 4 --no-cpu --no-ref-gpu > gpurun_out/r2_bench_8b_stream.log 2>&1; echo "bench 8b stream rc=$?"
grep '^{' gpurun_out/r2_bench_8b_stream.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['accept_len'], d['parity_check']['mismatches'], d['gpu_launches'])"
timeout 600 python tools/bench_attn.py > gpurun_out/r2_bench_attn.log 2>&1; echo "bench attn rc=$?"; cat gpurun_out/r2_bench_attn.log | grep -v Warn
