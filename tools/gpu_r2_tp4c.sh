#!/bin/bash
# 4 GPUs, final build: the driver's N=4 command (default steps / warmup)
mkdir -p gpurun_out
( time timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29814 bench.py --gpus 4 > gpurun_out/r2f7_bench_70b_tp4.log 2>&1 ) 2>&1 | tail -3; echo "bench tp4 rc=$?"
grep '^{' gpurun_out/r2f7_bench_70b_tp4.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['accept_len'], d['parity_check']['mismatches'], d['allreduce'], d['gpu_launches'], d['e2e']['value'])"
