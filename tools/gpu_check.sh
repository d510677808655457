#!/bin/bash
# Run the GPU parity tests group by group, each in its own process under a timeout, so a crash or a
# trapped kernel in one group does not hide the others.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
python -c "import ssd_b200.lib as l; l.load(); print('libssdk loaded')" > gpurun_out/load.log 2>&1
run() {
  name=$1; shift
  timeout 600 python -m pytest "$@" -q -x --no-header -p no:cacheprovider > gpurun_out/test_$name.log 2>&1
  echo "$name rc=$?" | tee -a gpurun_out/summary.txt
  tail -3 gpurun_out/test_$name.log
}
: > gpurun_out/summary.txt
run linear tests/test_ops_gpu.py -k "linear"
run gateup tests/test_ops_gpu.py -k "gate_up"
run norm_rope tests/test_ops_gpu.py -k "rms_norm or rope"
run attn tests/test_ops_gpu.py -k "paged_attention"
run sample tests/test_ops_gpu.py -k "sample"
run verify tests/test_ops_gpu.py -k "verify"
run engine tests/test_engine_gpu.py
cat gpurun_out/summary.txt
