"""The SOURCE of csrc/draft_persistent.cuh (experimental persistent draft forward) compiled for the host with
tests/emu/cuda_emu.h — one OS thread per CUDA thread, real barriers, real warp shuffles, the real device-wide barrier —
and run on three chained decode forwards, checked against the pinned oracle.  Complements the algorithm restatement of
test_draft_persistent_algo_cpu.py: this executes the kernel's own C++ (indexing, barrier placement, launch-counter
protocol); what it cannot show is device-only behaviour (memory model, L1 staleness, occupancy, speed)."""
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

from oracle.model import ModelCfg, OracleModel, random_weights

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "emu", "run_draft_persistent.cpp")
BIN = os.path.join(ROOT, "tests", "emu", "_build", "run_draft_persistent")

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++ (C++20)")
# SSD_B200_TSAN=1: build the emulated kernel with -fsanitize=thread; any unsynchronised conflicting access to "shared" or
# "global" memory (a missing __syncthreads, a vector read before the device-wide barrier) is then reported as a data race.
TSAN = os.environ.get("SSD_B200_TSAN") == "1"
if TSAN:
    BIN += "_tsan"


def _build():
    deps = [SRC, os.path.join(ROOT, "tests", "emu", "cuda_emu.h"), os.path.join(ROOT, "ssd_b200", "csrc", "draft_persistent.cuh"),
            os.path.join(ROOT, "ssd_b200", "csrc", "common.cuh")]
    if os.path.exists(BIN) and all(os.path.getmtime(BIN) >= os.path.getmtime(d) for d in deps):
        return
    os.makedirs(os.path.dirname(BIN), exist_ok=True)
    flags = ["-fsanitize=thread", "-g"] if TSAN else []
    subprocess.run(["g++", "-std=c++20", "-O1", "-pthread", "-Wno-unknown-pragmas", "-Wno-attributes", *flags, "-o", BIN, SRC],
                   check=True)


def _u16(t):
    return t.contiguous().view(torch.int16).numpy().astype(np.uint16)


@pytest.mark.parametrize("family,grid", [("llama", 3), ("qwen", 2)])
def test_persistent_draft_kernel_source_on_host_threads(tmp_path, family, grid):
    _build()
    torch.manual_seed(1)
    hd = 64 if family == "llama" else 128
    cfg = ModelCfg(hidden=256, layers=2, heads=4 if family == "llama" else 2, kv_heads=2 if family == "llama" else 1,
                   head_dim=hd, ffn=512, vocab=512, max_pos=256, rms_eps=1e-5 if family == "llama" else 1e-6,
                   rope_theta=500000.0, qk_norm=(family != "llama"))
    w = random_weights(cfg, seed=9)
    bs, nblk = 16, 6
    model = OracleModel(cfg, w, num_blocks=nblk, block_size=bs)
    bt = [4, 1, 5, 0, 3, 2]
    n = 21
    prompt = torch.randint(0, cfg.vocab, (n,))
    slots = torch.tensor([bt[p // bs] * bs + p % bs for p in range(n)], dtype=torch.int32)
    btt = torch.tensor([bt], dtype=torch.int32)
    model.forward(prompt, torch.arange(n), slots, torch.tensor([n], dtype=torch.int32), btt, n)
    kv0 = model.kv_cache.clone()

    # oracle: three chained greedy decode forwards
    toks, want = [77], []
    for step in range(3):
        p = n + step
        slot = torch.tensor([bt[p // bs] * bs + p % bs], dtype=torch.int32)
        hidden = model.forward(torch.tensor([toks[-1]]), torch.tensor([p]), slot, torch.tensor([p + 1], dtype=torch.int32), btt, 1)
        lg = model.compute_logits(hidden)[0].float().numpy()
        want.append(lg)
        toks.append(int(lg.argmax()))

    blob = tmp_path / "in.bin"
    with open(blob, "wb") as f:
        np.array([cfg.hidden, cfg.layers, cfg.heads, cfg.kv_heads, hd, cfg.ffn, cfg.vocab, int(cfg.qk_norm), bs, len(bt),
                  nblk * bs, n, 3, grid, cfg.max_pos], dtype=np.int32).tofile(f)
        np.array([cfg.rms_eps], dtype=np.float32).tofile(f)
        np.array(toks[:3], dtype=np.int64).tofile(f)
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
    nl = 3 * cfg.vocab
    got = torch.from_numpy(raw[:nl].astype(np.int16)).view(torch.bfloat16).float().numpy().reshape(3, cfg.vocab)
    ncache = cfg.layers * nblk * bs * cfg.kv_heads * hd
    kc = torch.from_numpy(raw[nl:nl + ncache].astype(np.int16)).view(torch.bfloat16).float().numpy()
    vc = torch.from_numpy(raw[nl + ncache:nl + 2 * ncache].astype(np.int16)).view(torch.bfloat16).float().numpy()
    for step in range(3):
        scale = np.abs(want[step]).max()
        err = np.abs(got[step] - want[step]).max()
        assert err <= 0.02 * scale + 0.02, (step, err, scale)
        assert int(got[step].argmax()) == int(want[step].argmax())
    ref = model.kv_cache.float().numpy()
    assert np.abs(kc - ref[0].reshape(-1)).max() <= 0.02 * np.abs(ref[0]).max() + 1e-3
    assert np.abs(vc - ref[1].reshape(-1)).max() <= 0.02 * np.abs(ref[1]).max() + 1e-3
