#!/bin/bash
# validation of the 256-token chunks (UMMA N = 128 / 256, verify batch 32) + prefill timing + a fresh launch list
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | grep -v "Warning\|warn" | tail -15 ) > gpurun_out/r2_pytest_gpu2.txt 2>&1; tail -15 gpurun_out/r2_pytest_gpu2.txt | cut -c1-300
timeout 600 python tools/bench_prefill.py > gpurun_out/r2_prefill.log 2>&1; echo "prefill rc=$?"; grep -v Warn gpurun_out/r2_prefill.log | tail -8
timeout 400 python bench.py --workload 8b --steps 24 --warmup 4 --no-cpu --no-ref-gpu > gpurun_out/r2_bench_8b_check.log 2>&1; echo "bench 8b rc=$?"
grep '^{' gpurun_out/r2_bench_8b_check.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['accept_len'], d['parity_check']['mismatches'])"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_70b.csv python tools/profile_step.py 70b > gpurun_out/r2_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
