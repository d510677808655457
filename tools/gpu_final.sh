#!/bin/bash
# what the driver runs at round end, on one GPU
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() { name=$1; shift; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" | tee -a gpurun_out/summary.txt; tail -3 gpurun_out/$name.log | cut -c1-1500; }
run pytest_gpu python -m pytest tests/ -x -q -m gpu
run smoke python -c "import __graft_entry__ as g; g.smoke()"
run bench_default python bench.py
run bench_reference python bench.py --impl reference --steps 2 --warmup 1
cat gpurun_out/summary.txt
