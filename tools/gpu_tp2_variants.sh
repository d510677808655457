#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() { name=$1; shift; timeout 240 "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" | tee -a gpurun_out/summary.txt; tail -4 gpurun_out/$name.log | cut -c1-900; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
run test_tp python -m pytest tests/test_tp_gpu.py -q -x --no-header -p no:cacheprovider
run bench_70b_tp2_symm $TR --master-port 29512 bench.py --gpus 2 --workload 70b --steps 16 --warmup 4 --no-cpu
SSD_B200_NO_SYMM=1 run bench_70b_tp2_nccl $TR --master-port 29513 bench.py --gpus 2 --workload 70b --steps 16 --warmup 4 --no-cpu
run trace_70b_tp2 $TR --master-port 29514 tools/trace_step.py 70b
tail -14 gpurun_out/trace_70b_tp2.log
cat gpurun_out/summary.txt
