#!/bin/bash
# 2 GPUs: TP=2 parity tests on the final build (both publish paths), the large-batch engine test, one 70B TP=2 line
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_tp_gpu.py "tests/test_engine_gpu.py::test_large_batch_spec_steps_match_oracle" -q --no-header -p no:cacheprovider 2>&1 | grep -v "Warning\|warn" | tail -12 ) > gpurun_out/r2_test_tp.txt; tail -8 gpurun_out/r2_test_tp.txt | cut -c1-300
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29811 bench.py --gpus 2 --workload 70b --steps 24 --warmup 4 --no-cpu > gpurun_out/r2_bench_70b_tp2.log 2>&1; echo "bench tp2 rc=$?"
grep '^{' gpurun_out/r2_bench_70b_tp2.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['accept_len'], d['parity_check']['mismatches'], d['allreduce'], d['gpu_launches'])"
