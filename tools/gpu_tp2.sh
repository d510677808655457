#!/bin/bash
# 2 GPUs: parity of the rewritten small kernels (1 GPU), the TP=2 tests (NCCL and one-shot all-reduce), the 70B TP=2 bench
mkdir -p gpurun_out
: > gpurun_out/summary_tp2.txt
run() { name=$1; shift; timeout 300 "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" | tee -a gpurun_out/summary_tp2.txt; tail -4 gpurun_out/$name.log; }
run test_ops python -m pytest tests/test_ops_gpu.py tests/test_engine_gpu.py -m gpu -q -x --no-header -p no:cacheprovider
run test_tp python -m pytest tests/test_tp_gpu.py -q -x --no-header -p no:cacheprovider
run bench_70b_tp2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload 70b --steps 16 --warmup 4 --no-cpu
grep '^{' gpurun_out/bench_70b_tp2.log > gpurun_out/bench_70b_tp2.json
cat gpurun_out/summary_tp2.txt
