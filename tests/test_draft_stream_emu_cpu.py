"""The SOURCE of csrc/draft_stream.cuh (the streaming draft kernel: K+1 draft forwards + in-kernel sampling in one
persistent launch) compiled for the host with tests/emu/cuda_emu.h — one OS thread per CUDA thread, real block barriers,
real warp shuffles, the real device-wide barrier, mbarrier / bulk-copy stand-ins — and checked against the pinned oracle:
logits of every forward, the KV rows written, the greedy tokens chained INSIDE the launch, and (temp > 0) the Philox
exponential-race tokens against oracle.verify.sample on the kernel's own logits.  This executes the kernel's own C++
(job sequence per CTA, ring slot / parity arithmetic, stage geometry for K <= 2048 / 4096 / 8192, barrier placement);
what it cannot show is device-only behaviour (async-proxy ordering, occupancy, speed)."""
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

from oracle.model import ModelCfg, OracleModel, random_weights

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "emu", "run_draft_stream.cpp")
BIN = os.path.join(ROOT, "tests", "emu", "_build", "run_draft_stream")

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++ (C++20)")
# SSD_B200_TSAN=1: build the emulated kernel with -fsanitize=thread; any unsynchronised conflicting access to "shared" or
# "global" memory (a missing __syncthreads, a vector read before the device-wide barrier) is then reported as a data race.
TSAN = os.environ.get("SSD_B200_TSAN") == "1"
if TSAN:
    BIN += "_tsan"


def _build():
    deps = [SRC, os.path.join(ROOT, "tests", "emu", "cuda_emu.h"), os.path.join(ROOT, "ssd_b200", "csrc", "draft_stream.cuh"),
            os.path.join(ROOT, "ssd_b200", "csrc", "common.cuh")]
    if os.path.exists(BIN) and all(os.path.getmtime(BIN) >= os.path.getmtime(d) for d in deps):
        return
    os.makedirs(os.path.dirname(BIN), exist_ok=True)
    flags = ["-fsanitize=thread", "-g"] if TSAN else []
    subprocess.run(["g++", "-std=c++20", "-O1", "-pthread", "-Wno-unknown-pragmas", "-Wno-attributes", *flags, "-o", BIN, SRC],
                   check=True)


def _u16(t):
    return t.contiguous().view(torch.int16).numpy().astype(np.uint16)


@pytest.mark.parametrize("family,grid,dims,temp,n_fwd,n,bs", [
    ("llama", 3, (256, 512), 0.0, 3, 21, 16),     # K = 256 / 512: 8 rows per job, one slot (gate|up: 8 pairs, two slots)
    ("qwen", 2, (256, 512), 0.8, 3, 270, 64),     # q/k norm, head_dim 128, Philox sampling in the kernel; context > 256:
                                                  # 16 KV splits per head (head_dim 128), partials + ticket + merge by the last split
    ("llama", 3, (256, 4096), 0.0, 2, 21, 16),    # down-proj K = 4096: 4 rows x 2 segments per job
    ("llama", 2, (256, 5120), 0.7, 2, 21, 16),    # down-proj K = 5120: 2 rows x 4 segments per job
    ("llama", 3, (256, 512), 0.0, 2, 1100, 64),   # 16 KV splits per head, 8 tokens per warp iteration, two load rounds per split,
                                                  # several attention units per CTA
    ("llama", 2, (256, 512), 0.0, 4, 254, 16),    # context 255 .. 258 inside ONE launch: the forwards switch from 8 KV splits
                                                  # (4 tokens per warp iteration) to 16 splits (8 tokens) at 257 tokens
])
def test_draft_stream_kernel_source_on_host_threads(tmp_path, family, grid, dims, temp, n_fwd, n, bs):
    from oracle import verify as V
    _build()
    torch.manual_seed(1)
    hd = 64 if family == "llama" else 128
    hidden, ffn = dims
    heads = hidden // hd if family == "llama" else max(2, hidden // hd)
    cfg = ModelCfg(hidden=hidden, layers=2, heads=heads, kv_heads=max(1, heads // 2), head_dim=hd, ffn=ffn, vocab=264,
                   max_pos=2048 if n > 1000 else 512, rms_eps=1e-5 if family == "llama" else 1e-6, rope_theta=500000.0, qk_norm=(family != "llama"))
    w = random_weights(cfg, seed=9)
    # page table: 6 entries (staged in shared memory by the kernel) or, for the K = 4096 case, 40 entries (> kDsBtSmem:
    # the kernel reads the table from global memory)
    nblk = 40 if ffn == 4096 else max(6, (n + n_fwd) // bs + 2)
    model = OracleModel(cfg, w, num_blocks=nblk, block_size=bs)
    bt = [4, 1, 5, 0, 3, 2] + list(range(6, nblk))
    prompt = torch.randint(0, cfg.vocab, (n,))
    slots = torch.tensor([bt[p // bs] * bs + p % bs for p in range(n)], dtype=torch.int32)
    btt = torch.tensor([bt], dtype=torch.int32)
    model.forward(prompt, torch.arange(n), slots, torch.tensor([n], dtype=torch.int32), btt, n)
    kv0 = model.kv_cache.clone()
    seed, call_base = 1234, 7 * 16

    blob = tmp_path / "in.bin"
    with open(blob, "wb") as f:
        np.array([cfg.hidden, cfg.layers, cfg.heads, cfg.kv_heads, hd, cfg.ffn, cfg.vocab, int(cfg.qk_norm), bs, len(bt),
                  nblk * bs, n, n_fwd, grid, cfg.max_pos, 3, 1], dtype=np.int32).tofile(f)
        np.array([cfg.rms_eps, temp], dtype=np.float32).tofile(f)
        np.array([seed, call_base], dtype=np.uint64).tofile(f)
        np.array([77], dtype=np.int64).tofile(f)
        np.array(bt, dtype=np.int32).tofile(f)
        for t in (w["embed"], w["final_norm"], w["lm_head"]):
            _u16(t).tofile(f)
        model.rope.numpy().astype(np.float32).tofile(f)
        ones = torch.ones(hd, dtype=torch.bfloat16)
        for lw in w["layers"]:
            for k in ("qkv", "o", "gate_up", "down", "input_norm", "post_norm"):
                _u16(lw[k]).tofile(f)
            _u16(lw.get("q_norm", ones)).tofile(f)
            _u16(lw.get("k_norm", ones)).tofile(f)
        _u16(kv0[0]).tofile(f)  # [L, nblk, bs, KV, hd] == [L, slots, KV, hd]
        _u16(kv0[1]).tofile(f)
    out = tmp_path / "out.bin"
    res = subprocess.run([BIN, str(blob), str(out)], capture_output=True, text=True, timeout=3000)
    assert res.returncode == 0, res.stderr[-2000:]
    assert "ThreadSanitizer" not in res.stderr, res.stderr[:3000]
    raw = np.fromfile(out, dtype=np.uint16)
    nl = n_fwd * cfg.vocab
    got = torch.from_numpy(raw[:nl].astype(np.int16)).view(torch.bfloat16).reshape(n_fwd, cfg.vocab)
    ncache = cfg.layers * nblk * bs * cfg.kv_heads * hd
    kc = torch.from_numpy(raw[nl:nl + ncache].astype(np.int16)).view(torch.bfloat16).float().numpy()
    vc = torch.from_numpy(raw[nl + ncache:nl + 2 * ncache].astype(np.int16)).view(torch.bfloat16).float().numpy()
    toks = np.frombuffer(raw[nl + 2 * ncache:].tobytes(), dtype=np.int64).tolist()
    assert toks[0] == 77 and len(toks) == n_fwd + 1

    # oracle: teacher-forced on the kernel's tokens; the last forward ran without lm_head (only its KV is checked)
    for step in range(n_fwd):
        p = n + step
        slot = torch.tensor([bt[p // bs] * bs + p % bs], dtype=torch.int32)
        hidden_o = model.forward(torch.tensor([toks[step]]), torch.tensor([p]), slot, torch.tensor([p + 1], dtype=torch.int32), btt, 1)
        if step == n_fwd - 1:
            assert toks[step + 1] == -1  # skip_last_head: nothing sampled
            break
        want = model.compute_logits(hidden_o)[0].float().numpy()
        g = got[step].float().numpy()
        scale = np.abs(want).max()
        assert np.abs(g - want).max() <= 0.02 * scale + 0.02, (step, np.abs(g - want).max(), scale)
        # the token the kernel sampled == the sampler's rule applied to the kernel's OWN bf16 logits
        mine = int(V.sample(got[step][None], torch.tensor([temp]), seed, call_base + step)[0])
        assert toks[step + 1] == mine, (step, toks[step + 1], mine)
    ref = model.kv_cache.float().numpy()
    assert np.abs(kc - ref[0].reshape(-1)).max() <= 0.02 * np.abs(ref[0]).max() + 1e-3
    assert np.abs(vc - ref[1].reshape(-1)).max() <= 0.02 * np.abs(ref[1]).max() + 1e-3
