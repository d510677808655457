#!/bin/bash
N=${N:-4}
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() { name=$1; shift; timeout 300 "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" | tee -a gpurun_out/summary.txt; tail -3 gpurun_out/$name.log | cut -c1-700; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run test_1gpu python -m pytest tests/test_ops_gpu.py tests/test_engine_gpu.py -q -x --no-header -p no:cacheprovider -k "linear or engine or spec or trace or resident or temperature"
run test_tp python -m pytest tests/test_tp_gpu.py -q -x --no-header -p no:cacheprovider
run bench_70b_tp$N $TR --master-port 29512 bench.py --gpus $N --workload 70b --steps 24 --warmup 4 --no-cpu
run trace_70b_tp$N $TR --master-port 29514 tools/trace_step.py 70b
grep -E "^gemm|^norm|^misc|^attn|^rope|target layer|forward|span" gpurun_out/trace_70b_tp$N.log | tail -12 | cut -c1-500
cat gpurun_out/summary.txt
