"""Workload for the per-kernel-family `ncu --set full` captures (tools/gpu_ncu_families.sh).

    python tools/ncu_workload.py engine   # eager (no CUDA graph) spec steps: every kernel family of the hot path at the
                                          # widths of Llama-3.1-70B (2 layers) + Llama-3.2-1B (2 layers), full vocabulary
    python tools/ncu_workload.py attn     # paged attention alone at the 70B verify shape (H=64, KV=8, hd=128, 7 queries) and
                                          # the 1B decode shape (H=32, KV=8, hd=64, 1 query), context 640 / 2048 / 8192
    python tools/ncu_workload.py gemm     # the projection GEMMs of one 70B layer at M=7 (qkv, o, gate|up + SiLU, down)
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
mode = sys.argv[1] if len(sys.argv) > 1 else "engine"
dev = torch.device("cuda:0")

if mode == "engine":
    import tempfile
    import random
    from ssd_b200 import lib as L, synth
    from ssd_b200.llm import LLM
    root = tempfile.mkdtemp()
    llm = LLM(synth.make_model_dir(root, "llama-3.1-70b", "target", layers=2), speculate=True,
              draft=synth.make_model_dir(root, "llama-3.2-1b", "draft", layers=2), speculate_k=6, num_gpus=1, max_num_seqs=1,
              max_model_len=4096, jit_speculate=True, enforce_eager=True, use_cuda_graph=False, use_pdl=False)
    r = llm.runner
    random.seed(0)
    prompt = [random.randint(0, 10000) for _ in range(640)]
    bt = list(range(r.max_blocks))
    rec = r.prefill(L.TARGET, prompt, bt)
    r.prefill(L.DRAFT, prompt, bt, want_sample=False)
    ctx = len(prompt)
    temp = float(os.environ.get("NCU_TEMP", "0"))
    for _ in range(3):
        toks, nacc, nrec = r.spec_step([ctx], [rec], [bt], [bt], [temp], [temp])
        ctx += int(nacc[0]) + 1
        rec = int(nrec[0])
    torch.cuda.synchronize()
elif mode == "attn":
    from ssd_b200 import ops
    for (H, KV, hd, q_len) in ((64, 8, 128, 7), (32, 8, 64, 1)):
        for ctx in (640, 2048, 8192):
            bs = 256
            nblk = (ctx + bs - 1) // bs
            kc = torch.randn(nblk, bs, KV, hd, device=dev).to(torch.bfloat16)
            vc = torch.randn(nblk, bs, KV, hd, device=dev).to(torch.bfloat16)
            bt = torch.arange(nblk, device=dev, dtype=torch.int32)[None].contiguous()
            q = torch.randn(q_len, H, hd, device=dev).to(torch.bfloat16)
            cl = torch.tensor([ctx], device=dev, dtype=torch.int32)
            for _ in range(2):
                ops.paged_attention(q, kc, vc, bt, cl, q_len, hd ** -0.5)
    torch.cuda.synchronize()
else:
    from ssd_b200 import ops
    d, ffn, qkv, M = 8192, 28672, 10240, 7
    x = torch.randn(M, d, device=dev).to(torch.bfloat16)
    xa = torch.randn(M, ffn, device=dev).to(torch.bfloat16)
    mats = {"qkv": (torch.randn(qkv, d, device=dev) * 0.02).to(torch.bfloat16), "o": (torch.randn(d, d, device=dev) * 0.02).to(torch.bfloat16),
            "down": (torch.randn(d, ffn, device=dev) * 0.02).to(torch.bfloat16)}
    gu = (torch.randn(2 * ffn, d, device=dev) * 0.02).to(torch.bfloat16)
    for _ in range(3):
        ops.linear(x, mats["qkv"])
        ops.linear(x, mats["o"])
        ops.gate_up_silu(x, gu)
        ops.linear(xa, mats["down"])
    torch.cuda.synchronize()
print("done")
