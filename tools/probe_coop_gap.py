"""Does the cooperative launch attribute of draft_stream_kernel cost time BETWEEN steps?  Same process, same weights:
resident-loop step time with SSDK_DRAFT_COOP=1 (default) and =0 (plain launch) on a 2-layer 8B-width target + the full 1B
draft (tools/check_draft_stream.py's configuration, where the loop measured 7.8 ms per step against a 5.4 ms step span)."""
import os
import random
import sys
import tempfile

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ssd_b200 import lib as L, synth  # noqa: E402
from ssd_b200.llm import LLM  # noqa: E402

root = tempfile.mkdtemp()
tdir = synth.make_model_dir(root, "llama-3.1-8b", "target", layers=2)
ddir = synth.make_model_dir(root, "llama-3.2-1b", "draft")
random.seed(0)
prompt = [random.randint(0, 10000) for _ in range(200)]
out = {}
for coop in ("1", "0", "1"):
    os.environ["SSDK_DRAFT_COOP"] = coop
    llm = LLM(tdir, speculate=True, draft=ddir, speculate_k=6, num_gpus=1, max_num_seqs=1, max_model_len=2048, jit_speculate=True)
    r = llm.runner
    bt = list(range(r.max_blocks))
    rec = r.prefill(L.TARGET, prompt, bt)
    r.prefill(L.DRAFT, prompt, bt, want_sample=False)
    r.stage([len(prompt)], [rec], [bt], [bt], [0.0], [0.0])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(4):
        r.step_resident(1)
    e0.record()
    for _ in range(24):
        r.step_resident(1)
    e1.record()
    torch.cuda.synchronize()
    toks, total, _ = r.fetch(1)
    print(f"SSDK_DRAFT_COOP={coop}: {e0.elapsed_time(e1) / 24:.3f} ms/step, tokens so far {int(total[0])}, last {toks[0].tolist()}", flush=True)
    llm.exit()
    del llm, r
    torch.cuda.empty_cache()
