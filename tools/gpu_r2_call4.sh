#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/check_draft_stream.py > gpurun_out/r2_stream_check.log 2>&1; echo "stream check rc=$?"; tail -3 gpurun_out/r2_stream_check.log | cut -c1-1200
timeout 900 python tools/check_draft_stream.py --temp 0.7 > gpurun_out/r2_stream_check_t07.log 2>&1; echo "stream check t0.7 rc=$?"; tail -2 gpurun_out/r2_stream_check_t07.log | cut -c1-1200
( timeout 900 python -m pytest tests/test_true_width_gpu.py -q -s 2>&1 | grep -v "Warning\|warn" | tail -60 ) > gpurun_out/r2_true_width.txt 2>&1; tail -40 gpurun_out/r2_true_width.txt
timeout 300 python tools/trace_step.py 70b:8 > gpurun_out/r2_trace_stream.log 2>&1; echo "trace rc=$?"; grep -v Warning gpurun_out/r2_trace_stream.log | grep "step span\|avg=\|^forward" | head -20
for wl in 8b 70b; do
  timeout 400 python bench.py --workload $wl --steps 24 --warmup 4 --no-cpu --no-ref-gpu > gpurun_out/r2_bench_${wl}_stream.log 2>&1; echo "bench $wl stream rc=$?"
  grep '^{' gpurun_out/r2_bench_${wl}_stream.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['accept_len'], d['parity_check']['mismatches'], d['gpu_launches'])"
done
for a in 4 16; do
  SSDK_DRAFT_L2_AHEAD=$a timeout 400 python bench.py --workload 8b --steps 24 --warmup 4 --no-cpu --no-ref-gpu > gpurun_out/r2_bench_8b_stream_l2_$a.log 2>&1; echo "bench 8b l2_ahead=$a rc=$?"
  grep '^{' gpurun_out/r2_bench_8b_stream_l2_$a.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['accept_len'])"
done
