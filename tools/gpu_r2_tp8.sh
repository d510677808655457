#!/bin/bash
# 8 GPUs: 70B TP=8 bench (default switches) + in-graph timeline
mkdir -p gpurun_out
run() { port=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $port "$@"; }
run 29711 bench.py --gpus 8 --workload 70b --steps 24 --warmup 4 --no-cpu > gpurun_out/r2_bench_70b_tp8.log 2>&1; echo "bench 70b tp8 rc=$?"
grep '^{' gpurun_out/r2_bench_70b_tp8.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['accept_len'], d['parity_check']['mismatches'], d['allreduce'], d['gpu_launches'], d['roofline']['step_frac'])"
run 29712 tools/trace_step.py 70b > gpurun_out/r2_timeline_70b_tp8.txt 2>&1; echo "trace tp8 rc=$?"
grep -v Warning gpurun_out/r2_timeline_70b_tp8.txt | grep "step span\|avg=\|layer sample\|^forward" | tail -14
