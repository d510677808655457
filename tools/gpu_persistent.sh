#!/bin/bash
# First hardware run of the EXPERIMENTAL persistent draft forward (csrc/draft_persistent.cuh): acceptance check against
# the regular path, then the step timeline and the 8B / 70B benches with it switched on.  Every command is under its own
# timeout: the kernel meets at device-wide barriers, a bug there traps after ~2 s instead of hanging the box.
mkdir -p gpurun_out
timeout 600 python tools/check_draft_persistent.py > gpurun_out/persistent_check.log 2>&1; echo "check rc=$?"; tail -3 gpurun_out/persistent_check.log
if [ -n "$SANITIZE" ]; then  # bring-up aid: out-of-bounds / misaligned accesses and shared-memory races of the persistent path
  for tool in memcheck racecheck; do
    SSDK_DRAFT_PERSISTENT=1 timeout 900 compute-sanitizer --tool $tool --log-file gpurun_out/persistent_$tool.log \
      python tools/check_draft_persistent.py --worker /tmp/persistent_$tool.npz > /dev/null 2>&1; echo "$tool rc=$?"; tail -4 gpurun_out/persistent_$tool.log
  done
fi
SSDK_DRAFT_PERSISTENT=1 timeout 200 python tools/trace_step.py 70b:8 > gpurun_out/trace_persistent.log 2>&1; echo "trace rc=$?"
grep -v Warning gpurun_out/trace_persistent.log | grep "step span\|avg=\|^forward"
for wl in 8b 70b; do
  SSDK_DRAFT_PERSISTENT=1 timeout 300 python bench.py --workload $wl --steps 16 --warmup 4 --no-cpu > gpurun_out/bench_${wl}_persistent.log 2>&1; echo "bench $wl rc=$?"
  grep '^{' gpurun_out/bench_${wl}_persistent.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['accept_len'])"
done
