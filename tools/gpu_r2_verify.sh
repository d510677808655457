#!/bin/bash
# final single-GPU verification of the committed build: the GPU test suite, smoke(), two short bench lines
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "Warning\|warn" | grep "passed\|failed\|FAILED\|Error\|\[KL\]\|\[true-width\|\[golden trace\|mean_accept_len" | tail -40 ) > gpurun_out/r2_pytest_gpu.txt 2>&1; cat gpurun_out/r2_pytest_gpu.txt | cut -c1-200 | tail -12
timeout 600 python __graft_entry__.py smoke > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_smoke.log
for wl in 8b 70b; do
timeout 400 python bench.py --workload $wl --steps 24 --warmup 4 --no-cpu --no-ref-gpu > gpurun_out/r2_bench_${wl}_verify.log 2>&1; echo "bench $wl rc=$?"
grep '^{' gpurun_out/r2_bench_${wl}_verify.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['accept_len'], d['parity_check']['mismatches'], d['gpu_launches'])"
done
