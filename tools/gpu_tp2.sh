#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() { name=$1; shift; timeout 240 "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" | tee -a gpurun_out/summary.txt; tail -4 gpurun_out/$name.log; }
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
run test_tp python -m pytest tests/test_tp_gpu.py -q -x --no-header -p no:cacheprovider
run bench_8b_tp2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload 8b --steps 24 --warmup 4 --no-cpu
run bench_70b_tp2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload 70b --steps 16 --warmup 4 --no-cpu
cat gpurun_out/summary.txt
