"""Prefill (TTFT) timing of ssdk_forward_tokens in chunks of 64 (round 1) vs 256 tokens (UMMA N = 256 instances of the GEMM):
a 2048-token prompt through the 8B-width target (full depth) and an 8-layer 70B-width target; then the reference bench's
prompt set (16 prompts x 128 tokens) one prompt per call vs packed two to a call (PairRunner.prefill_many)."""
import os
import random
import sys
import tempfile
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ssd_b200 import lib as L, synth  # noqa: E402
from ssd_b200.llm import LLM  # noqa: E402

for shape, layers in (("llama-3.1-8b", None), ("llama-3.1-70b", 8)):
    root = tempfile.mkdtemp()
    llm = LLM(synth.make_model_dir(root, shape, "target", layers=layers), speculate=True,
              draft=synth.make_model_dir(root, "llama-3.2-1b", "draft", layers=2), speculate_k=6, num_gpus=1, max_num_seqs=1,
              max_model_len=4096, jit_speculate=True)
    r = llm.runner
    random.seed(0)
    prompt = [random.randint(0, 10000) for _ in range(2048)]
    bt = list(range(r.max_blocks))
    out = {}
    for chunk in (64, 256):
        r.prefill(L.TARGET, prompt, bt, chunk=chunk)  # warm-up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        tok = r.prefill(L.TARGET, prompt, bt, chunk=chunk)
        torch.cuda.synchronize()
        out[chunk] = (time.perf_counter() - t0, tok)
    L_ = layers or synth.SHAPES[shape][1]
    print(f"{shape} ({L_} layers) 2048-token prefill: chunk 64 {out[64][0] * 1e3:.1f} ms, chunk 256 {out[256][0] * 1e3:.1f} ms "
          f"({out[64][0] / out[256][0]:.2f}x), {2048 / out[256][0]:.0f} tok/s, first token equal: {out[64][1] == out[256][1]}", flush=True)
    llm.exit()
    del llm, r
    torch.cuda.empty_cache()

# 16 x 128-token prompts (bench/bench.py --random: numseqs 16, input_len 128): one prompt per call vs prefill_many
root = tempfile.mkdtemp()
llm = LLM(synth.make_model_dir(root, "llama-3.1-8b", "target"), speculate=True,
          draft=synth.make_model_dir(root, "llama-3.2-1b", "draft", layers=2), speculate_k=6, num_gpus=1, max_num_seqs=16,
          max_model_len=1024, jit_speculate=True)
r = llm.runner
random.seed(1)
prompts = [[random.randint(0, 10000) for _ in range(128)] for _ in range(16)]
bts = [list(range(b * r.max_blocks, (b + 1) * r.max_blocks)) for b in range(16)]
res = {}
for mode in ("one by one", "packed"):
    for rep in range(2):  # first pass = warm-up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if mode == "packed":
            toks = r.prefill_many(L.TARGET, prompts, bts, [0] * 16)
        else:
            toks = [r.prefill(L.TARGET, prompts[b], bts[b]) for b in range(16)]
        torch.cuda.synchronize()
        res[mode] = (time.perf_counter() - t0, toks)
a, b = res["one by one"], res["packed"]
print(f"llama-3.1-8b, 16 prompts x 128 tokens: one per call {a[0] * 1e3:.1f} ms, packed {b[0] * 1e3:.1f} ms ({a[0] / b[0]:.2f}x), "
      f"{16 * 128 / b[0]:.0f} prompt tok/s, first tokens equal: {a[1] == b[1]}", flush=True)
llm.exit()
