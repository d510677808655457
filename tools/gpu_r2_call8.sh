#!/bin/bash
mkdir -p gpurun_out
SSDK_LIB=$PWD/ssd_b200/_lib/libssdk_trace.so timeout 300 python tools/trace_step.py 70b:8 > gpurun_out/r2_trace_stream_probe.log 2>&1; echo "trace probe rc=$?"; grep -v Warning gpurun_out/r2_trace_stream_probe.log | grep -A17 "draft_stream_kernel, second" | head -20
timeout 300 python tools/trace_step.py 70b:8 > gpurun_out/r2_trace_stream5.log 2>&1; echo "trace rc=$?"; grep "avg=" gpurun_out/r2_trace_stream5.log | head -2
timeout 900 python tools/check_draft_stream.py > gpurun_out/r2_stream_check.log 2>&1; echo "stream check rc=$?"; tail -1 gpurun_out/r2_stream_check.log | cut -c1-900
for wl in 8b 70b; do
timeout 400 python bench.py --workload $wl --steps 24 --warmup 4 --no-cpu --no-ref-gpu > gpurun_out/r2_bench_${wl}_stream.log 2>&1; echo "bench $wl stream rc=$?"
grep '^{' gpurun_out/r2_bench_${wl}_stream.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['accept_len'], d['parity_check']['mismatches'], d['gpu_launches'])"
done
