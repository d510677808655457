#!/bin/bash
N=8
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() { name=$1; shift; timeout 300 "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" | tee -a gpurun_out/summary.txt; tail -3 gpurun_out/$name.log | cut -c1-700; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run bench_70b_tp$N $TR --master-port 29512 bench.py --gpus $N --workload 70b --steps 24 --warmup 4 --no-cpu
run trace_70b_tp$N $TR --master-port 29514 tools/trace_step.py 70b
grep -E "^gemm|^norm|^misc|^attn|^rope|target layer|forward|span" gpurun_out/trace_70b_tp$N.log | tail -12 | cut -c1-500
run bench_qwen_tp4 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29516 bench.py --gpus 4 --workload qwen32b --steps 24 --warmup 4 --no-cpu
cat gpurun_out/summary.txt
