#!/bin/bash
# round 2, last GPU seconds: streaming draft kernel with 16 KV splits / 8 tokens per warp iteration beyond 1024 tokens —
# long context (prompt 3000) against the kernel-per-op path (tokens, accept lengths, logits, step time of both), then the
# short-context pair (correctness; the two modes run at the same time)
mkdir -p gpurun_out
timeout 70 python tools/check_draft_stream.py --prompt-len 3000 > gpurun_out/r2f7_stream_long.txt 2>&1; echo "long rc=$?"; tail -2 gpurun_out/r2f7_stream_long.txt | cut -c1-600
timeout 40 python tools/check_draft_stream.py --parallel > gpurun_out/r2f7_stream_short.txt 2>&1; echo "short rc=$?"; tail -2 gpurun_out/r2f7_stream_short.txt | cut -c1-600
