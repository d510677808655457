#!/bin/bash
mkdir -p gpurun_out
timeout 45 python tools/probe_coop_gap.py > gpurun_out/r2f9_coop_gap.txt 2>&1; echo rc=$?; grep COOP gpurun_out/r2f9_coop_gap.txt; tail -3 gpurun_out/r2f9_coop_gap.txt | cut -c1-300
