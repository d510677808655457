#!/bin/bash
# round 2, last call: the GPU suite on the final build (packed prefill test included), then the draft-stream A/B:
# norm operands requested before the barrier (now unconditional) and the idle-only L2 window (SSDK_DRAFT_L2_AHEAD = 0/4/8/16).
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2f4_pytest_gpu.txt 2>&1 ) 2>&1 | tail -3
tail -4 gpurun_out/r2f4_pytest_gpu.txt
for a in 0 4 8 16; do
  SSDK_DRAFT_L2_AHEAD=$a timeout 400 python bench.py --workload 8b --steps 48 --warmup 6 --no-ref-gpu --no-cpu \
    > gpurun_out/r2f4_bench_8b_l2a$a.json 2> gpurun_out/r2f4_bench_8b_l2a$a.err
  echo "8b l2_ahead=$a rc=$?"; python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r2f4_bench_8b_l2a$a.json") if l.startswith("{")][-1])
print({k: d.get(k) for k in ("ms_per_step", "value", "accept_len", "parity_check", "draft_path")})
PY
done
for a in 0 8; do
  SSDK_DRAFT_L2_AHEAD=$a timeout 300 python tools/trace_step.py 70b:8 > gpurun_out/r2f4_timeline_l2a$a.txt 2>&1
  echo "trace l2_ahead=$a rc=$?"; grep "per layer\|^misc" gpurun_out/r2f4_timeline_l2a$a.txt | head -5
done
for a in 0 8; do
  SSDK_DRAFT_L2_AHEAD=$a timeout 500 python bench.py --steps 48 --warmup 6 --no-ref-gpu --no-cpu \
    > gpurun_out/r2f4_bench_70b_l2a$a.json 2> gpurun_out/r2f4_bench_70b_l2a$a.err
  echo "70b l2_ahead=$a rc=$?"; cut -c1-200 gpurun_out/r2f4_bench_70b_l2a$a.json | tail -1
done
