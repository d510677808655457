#!/bin/bash
# 4 GPUs: 70B TP=4 A/B of the GEMM-fused all-reduce publish, Qwen3-32B + 0.6B at temp 0.7 (BASELINE config 4), timeline
mkdir -p gpurun_out
run() { torchrun_port=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $torchrun_port "$@"; }
for fp in 1 0; do
  SSDK_FUSED_PUBLISH=$fp run 2961$fp bench.py --gpus 4 --workload 70b --steps 24 --warmup 4 --no-cpu > gpurun_out/r2_bench_70b_tp4_fp$fp.log 2>&1; echo "bench 70b tp4 fused_publish=$fp rc=$?"
  grep '^{' gpurun_out/r2_bench_70b_tp4_fp$fp.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['accept_len'], d['parity_check']['mismatches'], d['allreduce'], d['gpu_launches'])"
done
for fp in 1 0; do
SSDK_FUSED_PUBLISH=$fp run 2963$fp bench.py --gpus 4 --workload qwen32b --temp 0.7 --lm-scale 10 --steps 24 --warmup 4 --no-cpu > gpurun_out/r2_bench_qwen32b_tp4_t07_fp$fp.log 2>&1; echo "bench qwen tp4 t0.7 fp=$fp rc=$?"
grep '^{' gpurun_out/r2_bench_qwen32b_tp4_t07_fp$fp.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['accept_len'], d['allreduce'], d['roofline']['step_frac'])"
done
run 29640 bench.py --gpus 4 --workload qwen32b --steps 24 --warmup 4 --no-cpu > gpurun_out/r2_bench_qwen32b_tp4_t0.log 2>&1; echo "bench qwen tp4 t0 rc=$?"
grep '^{' gpurun_out/r2_bench_qwen32b_tp4_t0.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['accept_len'], d['parity_check']['mismatches'], d['roofline']['step_frac'])"
run 29650 tools/trace_step.py 70b > gpurun_out/r2_timeline_70b_tp4.txt 2>&1; echo "trace tp4 rc=$?"
grep -v Warning gpurun_out/r2_timeline_70b_tp4.txt | grep "step span\|avg=\|layer sample\|^forward" | tail -14
