"""Stand-alone timing of ssdk_paged_attn (attention + split-KV combine) at the shapes of the speculative step.
Launches are captured in a CUDA graph (100 per replay) so the number is the kernel pair's GPU time without host launch
gaps; this is the warm-L2 figure (as in the step, where the new tokens' K/V were just written)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ssd_b200 import ops  # noqa: E402

dev = torch.device("cuda")
torch.manual_seed(0)
SHAPES = [  # name, H, KV, hd, q_len
    ("70b verify tp1", 64, 8, 128, 7), ("70b verify tp8", 8, 1, 128, 7), ("8b verify", 32, 8, 128, 7),
    ("qwen32b verify tp4", 16, 2, 128, 7), ("1b decode", 32, 8, 64, 1), ("0.6b decode", 16, 8, 128, 1),
]
bs, nblk = 256, 16


def time_shape(H, KV, hd, Q):
    kc = torch.randn(nblk, bs, KV, hd, device=dev).bfloat16()
    vc = torch.randn(nblk, bs, KV, hd, device=dev).bfloat16()
    bt = torch.arange(nblk, dtype=torch.int32, device=dev)[None]
    q = torch.randn(Q, H, hd, device=dev).bfloat16()
    row = []
    for ctx in (135, 400, 640, 2048):
        cl = torch.tensor([ctx], dtype=torch.int32, device=dev)
        ops.paged_attention(q, kc, vc, bt, cl, Q, hd ** -0.5)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            with torch.cuda.graph(g, stream=s):
                for _ in range(100):
                    ops.paged_attention(q, kc, vc, bt, cl, Q, hd ** -0.5)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        row.append(f"ctx {ctx}: {e0.elapsed_time(e1) * 1000 / 500:6.2f} us")
    return "  ".join(row)


for name, H, KV, hd, Q in SHAPES:
    print(f"{name:20s} " + time_shape(H, KV, hd, Q), flush=True)
