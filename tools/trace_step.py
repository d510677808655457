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
workload = sys.argv[1] if len(sys.argv) > 1 else "8b"   # "70b:8" = 70B dimensions with 8 layers (quick to set up)
workload, _, nl = workload.partition(":")
shapes = {"8b": ("llama-3.1-8b", "llama-3.2-1b"), "70b": ("llama-3.1-70b", "llama-3.2-1b")}[workload]
root = tempfile.mkdtemp()
llm = LLM(synth.make_model_dir(root, shapes[0], "target", layers=int(nl) if nl else None), speculate=True, draft=synth.make_model_dir(root, shapes[1], "draft"),
          speculate_k=6, num_gpus=world, max_num_seqs=1, max_model_len=int(sys.argv[sys.argv.index('--max-len') + 1]) if '--max-len' in sys.argv else 4096, jit_speculate=True, use_pdl=("--no-pdl" not in sys.argv))
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
cap = 32768
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
order = sorted(range(n), key=lambda i: int(t[i, 1]))
all_ids, all_ts = [int(t[i, 0]) for i in order], [int(t[i, 1]) for i in order]
# kernel-start marks (id < 16) and the phase marks (id >= 16) that follow each of them
kern = [i for i in range(n) if all_ids[i] < 16]
ids, ts = [all_ids[i] for i in kern], [all_ts[i] for i in kern]
nk = len(ids)
print("kernels:", nk, "marks:", n, "step span us:", (ts[-1] - ts[0]) / 1e3)
agg = collections.defaultdict(lambda: [0, 0.0])
for i in range(nk - 1):
    agg[names[ids[i]]][0] += 1
    agg[names[ids[i]]][1] += (ts[i + 1] - ts[i]) / 1e3
for k, (c, us) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k:8s} n={c:5d} total={us:9.1f} us avg={us / c:6.2f}")
seq = [(names[ids[i]], round((ts[i + 1] - ts[i]) / 1e3, 2)) for i in range(nk - 1)]
print("draft layer sample:", seq[10:20])
print("target layer sample:", seq[-40:-30])
preps = [i for i in range(nk) if ids[i] == 1]
for a, b in zip(preps, preps[1:] + [nk - 1]):
    print("forward", (ts[b] - ts[a]) / 1e3, "us", b - a, "kernels")


def phases(lo, hi, label):
    """mean gaps start -> phase marks -> next kernel start, per kernel kind, for kernels lo..hi-1"""
    acc = collections.defaultdict(lambda: collections.defaultdict(lambda: [0, 0.0]))
    for j in range(lo, min(hi, nk - 1)):
        a, b = kern[j], kern[j + 1]
        pts = [("start", all_ts[a])] + [(f"p{all_ids[i] % 8}", all_ts[i]) for i in range(a + 1, b)] + [("next", all_ts[b])]
        kind = names[ids[j]]
        if kind == "attn" and any(40 <= all_ids[i] < 48 for i in range(a + 1, b)):
            kind = "combine"
        for (n0, t0), (n1, t1) in zip(pts, pts[1:]):
            e = acc[kind][f"{n0}->{n1}"]
            e[0] += 1
            e[1] += (t1 - t0) / 1e3
    print(f"--- phase gaps (us), {label}")
    for kind, d in acc.items():
        print(f"  {kind:8s} " + "  ".join(f"{k}={v[1] / v[0]:.2f}" for k, v in d.items()))


# streaming draft kernel: phase points of its second forward (ids 64 + point, draft_stream.cuh: ds_mark)
pts = sorted((all_ts[i], all_ids[i] - 64) for i in range(n) if all_ids[i] >= 64)
if pts:
    label = {0: "A prologue done", 1: "A gemv done", 2: "A barrier", 3: "B attention done", 4: "B barrier", 5: "C gemv done",
             6: "C barrier", 7: "D prologue done", 8: "D gemv done", 9: "D barrier", 10: "E gemv done", 11: "E barrier",
             12: "probe barrier"}
    gaps = collections.defaultdict(lambda: [0, 0.0])
    for (t0, a), (t1, b) in zip(pts, pts[1:]):
        g = gaps[(a, b)]
        g[0] += 1
        g[1] += (t1 - t0) / 1e3
    print("--- draft_stream_kernel, second forward of the launch, mean gap (us) between phase points over the layers")
    tot = 0.0
    for (a, b), (c, us) in sorted(gaps.items()):
        print(f"  {label.get(a, a):>18s} -> {label.get(b, b):<18s} n={c:3d} {us / c:7.2f}")
        tot += us / c
    print(f"  per layer {tot:.1f} us; forward span {(pts[-1][0] - pts[0][0]) / 1e3:.1f} us")
if len(preps) >= 2:
    phases(preps[0], preps[1], "first draft forward")
if preps:
    phases(preps[-1], nk, "target verify forward")
