"""Real (PDL-overlapped, in-graph) per-kernel timeline of one speculative step via ssdk_debug_trace."""
import collections
import os
import random
import sys
import tempfile

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ssd_b200 import lib as L, synth  # noqa: E402
from ssd_b200.llm import LLM  # noqa: E402

import torch.distributed as dist  # noqa: E402

world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
workload = sys.argv[1] if len(sys.argv) > 1 else "8b"
shapes = {"8b": ("llama-3.1-8b", "llama-3.2-1b"), "70b": ("llama-3.1-70b", "llama-3.2-1b")}[workload]
root = tempfile.mkdtemp()
llm = LLM(synth.make_model_dir(root, shapes[0], "target"), speculate=True, draft=synth.make_model_dir(root, shapes[1], "draft"),
          speculate_k=6, num_gpus=world, max_num_seqs=1, max_model_len=4096, jit_speculate=True, use_pdl=("--no-pdl" not in sys.argv))
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
cap = 4096
buf = torch.zeros(cap, 2, dtype=torch.int64, device="cuda")
if world > 1:
    dist.barrier()
L.check(r.lib.ssdk_debug_trace(buf.data_ptr(), cap))
r.step_resident(1)
torch.cuda.synchronize()
L.check(r.lib.ssdk_debug_trace(None, 0))
if rank != 0:
    sys.exit(0)
t = buf.cpu()
n = int((t[:, 0] != 0).sum())
names = {1: "prep", 2: "norm", 3: "gemm", 4: "rope", 5: "attn", 6: "sample", 7: "verify", 8: "misc"}
ids, ts = t[:n, 0].tolist(), t[:n, 1].tolist()
print("marks:", n, "step span us:", (ts[-1] - ts[0]) / 1e3)
# increment attributed to kernel i = time from its dependency-resolved mark to the next kernel's mark
agg = collections.defaultdict(lambda: [0, 0.0])
for i in range(n - 1):
    agg[names[ids[i]]][0] += 1
    agg[names[ids[i]]][1] += (ts[i + 1] - ts[i]) / 1e3
for k, (c, us) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k:8s} n={c:5d} total={us:9.1f} us avg={us / c:6.2f}")
# one draft layer in detail (skip the first forward's prologue)
seq = [(names[ids[i]], round((ts[i + 1] - ts[i]) / 1e3, 2)) for i in range(n - 1)]
print("draft layer sample:", seq[10:20])
print("target layer sample:", seq[-40:-30])
# forward boundaries: prep marks
preps = [i for i in range(n) if ids[i] == 1]
for a, b in zip(preps, preps[1:] + [n - 1]):
    print("forward", (ts[b] - ts[a]) / 1e3, "us", b - a, "kernels")
