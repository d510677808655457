#!/bin/bash
# 2 GPUs, final build: the driver's N=2 command (default steps / warmup)
mkdir -p gpurun_out
( time timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29813 bench.py --gpus 2 > gpurun_out/r2f7_bench_70b_tp2.log 2>&1 ) 2>&1 | tail -3; echo "bench tp2 rc=$?"
grep '^{' gpurun_out/r2f7_bench_70b_tp2.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['accept_len'], d['parity_check']['mismatches'], d['allreduce'], d['gpu_launches'], d['e2e']['value'])"
