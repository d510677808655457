#!/bin/bash
# round-2 single-GPU measurement set for profiles/: default bench (with the reference GPU arm and the CPU arm), the other 1-GPU
# configs, the reference CPU arm alone, the ncu launch list of one step, the GEMM micro-bench, smoke() and the GPU test suite
mkdir -p gpurun_out
( time timeout 1200 python bench.py > gpurun_out/r2_bench_70b_tp1.json 2> gpurun_out/r2_bench_70b_tp1.err ) 2>&1 | tail -3; echo "bench default rc=$?"; cut -c1-400 gpurun_out/r2_bench_70b_tp1.json
( time timeout 900 python bench.py --workload 8b --steps 48 --warmup 6 > gpurun_out/r2_bench_8b_tp1.json 2> gpurun_out/r2_bench_8b_tp1.err ) 2>&1 | tail -3; cut -c1-300 gpurun_out/r2_bench_8b_tp1.json
( time timeout 900 python bench.py --workload qwen32b --temp 0.7 --lm-scale 10 --steps 48 --warmup 6 --no-cpu > gpurun_out/r2_bench_qwen32b_tp1_t07.json 2> gpurun_out/r2_bench_qwen32b_tp1_t07.err ) 2>&1 | tail -3; cut -c1-300 gpurun_out/r2_bench_qwen32b_tp1_t07.json
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err ) 2>&1 | tail -3; cut -c1-300 gpurun_out/r2_bench_reference_arm.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_70b.csv python tools/profile_step.py 70b > gpurun_out/r2_ncu_launches.log 2>&1; echo "ncu launches rc=$?"; tail -2 gpurun_out/r2_ncu_launches.log
timeout 600 python tools/bench_gemm.py > gpurun_out/r2_gemm_bench.log 2>&1; echo "gemm bench rc=$?"
timeout 600 python __graft_entry__.py smoke > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2_smoke.log
( time timeout 1800 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "Warning\|warn" | grep "passed\|failed\|FAILED\|Error\|\[KL\]\|\[true-width\|\[golden trace\|mean_accept_len" | tail -40 ) > gpurun_out/r2_pytest_gpu.txt 2>&1; cat gpurun_out/r2_pytest_gpu.txt | cut -c1-300
