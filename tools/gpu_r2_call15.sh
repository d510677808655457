#!/bin/bash
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "Warning\|warn" | grep "passed\|failed\|FAILED\|Error\|\[KL\]\|\[true-width\|\[golden trace\|mean_accept_len" | tail -40 ) > gpurun_out/r2_pytest_gpu.txt 2>&1; cat gpurun_out/r2_pytest_gpu.txt | cut -c1-260
