#!/bin/bash
# 2 GPUs: TP=2 parity tests (NCCL / one-shot all-reduce, fused publish, split-K SiLU), 70B TP=2 bench A/B
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_tp_gpu.py -q -x --no-header -p no:cacheprovider 2>&1 | grep -v "Warning\|warn" | tail -30 ) > gpurun_out/r2_test_tp.txt; tail -12 gpurun_out/r2_test_tp.txt
for fp in 1 0; do
  SSDK_FUSED_PUBLISH=$fp timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$fp bench.py --gpus 2 --workload 70b --steps 24 --warmup 4 --no-cpu > gpurun_out/r2_bench_70b_tp2_fp$fp.log 2>&1; echo "bench tp2 fused_publish=$fp rc=$?"
  grep '^{' gpurun_out/r2_bench_70b_tp2_fp$fp.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['accept_len'], d['parity_check'], d['allreduce'], d['gpu_launches'])"
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/trace_step.py 70b:8 > gpurun_out/r2_trace_70b_tp2.log 2>&1; echo "trace tp2 rc=$?"
grep -v Warning gpurun_out/r2_trace_70b_tp2.log | grep "step span\|avg=\|layer sample\|^forward" | head -24
for a in 0 8 24; do
  SSDK_DRAFT_L2_AHEAD=$a timeout 300 python bench.py --workload 8b --steps 24 --warmup 4 --no-cpu --no-ref-gpu > gpurun_out/r2_bench_8b_l2_$a.log 2>&1; echo "bench 8b l2_ahead=$a rc=$?"
  grep '^{' gpurun_out/r2_bench_8b_l2_$a.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['accept_len'])"
done
