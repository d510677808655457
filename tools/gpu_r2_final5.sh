#!/bin/bash
# round 2: draft-stream attention phase with the page table in shared memory and the first K/V loads requested before the
# q|k|v rebuild — GPU suite, then the same measurements as gpu_r2_final4.sh (8B + 1B: 8.483 ms, 40.8 us per draft layer before)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2f5_pytest_gpu.txt 2>&1 ) 2>&1 | tail -3
tail -3 gpurun_out/r2f5_pytest_gpu.txt
show() { python - "$1" <<'PY'
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
print({k: d.get(k) for k in ("ms_per_step", "value", "accept_len", "draft_path")}, d["parity_check"]["mismatches"], d["e2e"]["value"])
PY
}
timeout 400 python bench.py --workload 8b --steps 48 --warmup 6 --no-ref-gpu --no-cpu > gpurun_out/r2f5_bench_8b.json 2> gpurun_out/r2f5_bench_8b.err
echo "8b rc=$?"; show gpurun_out/r2f5_bench_8b.json
timeout 300 python tools/trace_step.py 70b:8 > gpurun_out/r2f5_timeline.txt 2>&1
echo "trace rc=$?"; grep "per layer\|^misc" gpurun_out/r2f5_timeline.txt; grep -A14 "draft_stream_kernel, second forward" gpurun_out/r2f5_timeline.txt | tail -14
timeout 500 python bench.py --steps 48 --warmup 6 --no-ref-gpu --no-cpu > gpurun_out/r2f5_bench_70b.json 2> gpurun_out/r2f5_bench_70b.err
echo "70b rc=$?"; show gpurun_out/r2f5_bench_70b.json
timeout 500 python bench.py --workload qwen32b --temp 0.7 --lm-scale 10 --steps 48 --warmup 6 --no-ref-gpu --no-cpu > gpurun_out/r2f5_bench_qwen.json 2> gpurun_out/r2f5_bench_qwen.err
echo "qwen rc=$?"; show gpurun_out/r2f5_bench_qwen.json
