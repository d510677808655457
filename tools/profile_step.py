"""Run exactly one device-resident speculative step between cudaProfilerStart/Stop (for ncu --profile-from-start off)."""
import os
import random
import sys
import tempfile

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ssd_b200 import lib as L, synth  # noqa: E402
from ssd_b200.llm import LLM  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "8b"
shapes = {"8b": ("llama-3.1-8b", "llama-3.2-1b"), "70b": ("llama-3.1-70b", "llama-3.2-1b"),
          "tiny": ("llama-tiny-target", "llama-tiny-draft")}[workload]
root = tempfile.mkdtemp()
llm = LLM(synth.make_model_dir(root, shapes[0], "target"), speculate=True, draft=synth.make_model_dir(root, shapes[1], "draft"),
          speculate_k=6, max_num_seqs=1, max_model_len=4096, jit_speculate=True, use_pdl=("--no-pdl" not in sys.argv))
r = llm.runner
random.seed(0)
prompt = [random.randint(0, 10000) for _ in range(128)]
bt = list(range(r.max_blocks))
rec = r.prefill(L.TARGET, prompt, bt)
r.prefill(L.DRAFT, prompt, bt, want_sample=False)
r.stage([len(prompt)], [rec], [bt], [bt], [0.0], [0.0])
for _ in range(3):
    r.step_resident(1)
torch.cuda.synchronize()
n0 = r.launch_count
torch.cuda.cudart().cudaProfilerStart()
r.step_resident(1)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled one step; launches in the step:", r.launch_count - n0)
