#!/bin/bash
mkdir -p gpurun_out
SSDK_PUBLISH_PER_PEER=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus 8 --workload 70b --steps 24 --warmup 4 --no-cpu > gpurun_out/r2_bench_70b_tp8_pp1.log 2>&1; echo "bench 70b tp8 per-peer publish rc=$?"
grep '^{' gpurun_out/r2_bench_70b_tp8_pp1.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['accept_len'], d['parity_check']['mismatches'], d['allreduce'], d['roofline']['step_frac'])"
