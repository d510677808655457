"""add_rmsnorm_kernel (split-K, dense and embedding inputs; 1-slice, 2-slice and staged instantiations),
rope_store_kernel (Llama and Qwen head-norm variants, KV scatter with a skipped slot) and silu_mul_kernel executed from
SOURCE on host threads (tests/emu) against the pinned oracle ops — the CPU counterpart of tests/test_ops_gpu.py, so that
a refactor of these kernels is checked before it costs GPU minutes."""
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

from oracle import ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "emu", "run_elementwise.cpp")
BIN = os.path.join(ROOT, "tests", "emu", "_build", "run_elementwise")
BF = torch.bfloat16
pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++ (C++20)")


def _build():
    deps = [SRC, os.path.join(ROOT, "tests", "emu", "cuda_emu.h"), os.path.join(ROOT, "ssd_b200", "csrc", "elementwise.cuh"),
            os.path.join(ROOT, "ssd_b200", "csrc", "common.cuh")]
    if os.path.exists(BIN) and all(os.path.getmtime(BIN) >= os.path.getmtime(d) for d in deps):
        return
    os.makedirs(os.path.dirname(BIN), exist_ok=True)
    subprocess.run(["g++", "-std=c++20", "-O1", "-pthread", "-Wno-unknown-pragmas", "-Wno-attributes", "-o", BIN, SRC], check=True)


def _u16(t):
    return t.contiguous().view(torch.int16).numpy().astype(np.uint16)


def _bf(a, shape):
    return torch.from_numpy(a.astype(np.int16)).view(BF).reshape(shape)


def _run(mode, parts, tmp_path):
    _build()
    inp, out = tmp_path / f"{mode}.in", tmp_path / f"{mode}.out"
    with open(inp, "wb") as f:
        for a in parts:
            np.ascontiguousarray(a).tofile(f)
    res = subprocess.run([BIN, mode, str(inp), str(out)], capture_output=True, text=True, timeout=1800)
    assert res.returncode == 0, res.stderr[-2000:]
    return np.fromfile(out, dtype=np.uint16)


def _ulp_frac(a, b):
    return (a.view(torch.int16) != b.view(torch.int16)).float().mean().item()


@pytest.mark.parametrize("M,d,S,threads,use_ids,has_res", [
    (3, 256, 8, 32, 0, 1),     # draft-like: one slice, 8 split-K slabs
    (2, 8192, 4, 512, 0, 1),   # 70B-like: two slices, 4 slabs
    (2, 1024, 0, 128, 0, 1),   # dense bf16 input
    (3, 512, 0, 64, 1, 0),     # embedding gather, first layer (no residual in)
    (1, 640, 3, 32, 0, 1),     # staged (SLICES = 0) path: 640 / (8 * 32) = 2.5 slices
])
def test_norm_source(tmp_path, M, d, S, threads, use_ids, has_res):
    g = torch.Generator().manual_seed(d + S)
    vocab, eps = 50, 1e-5
    w = (1 + 0.1 * torch.randn(d, generator=g)).to(BF)
    residual = torch.randn(M, d, generator=g).to(BF)
    partial = torch.randn(max(S, 1), M, d, generator=g)
    dense = torch.randn(M, d, generator=g).to(BF)
    ids = torch.tensor([7, 49, 0][:M] + [3] * max(0, M - 3), dtype=torch.int64)
    embed = torch.randn(vocab, d, generator=g).to(BF)
    parts = [np.array([M, d, S, threads, use_ids, vocab, has_res], dtype=np.int32), np.array([eps], dtype=np.float32), _u16(w),
             _u16(residual)]
    if S > 0:
        parts.append(partial[:S].numpy().astype(np.float32))
    else:
        parts.append(_u16(dense))
    parts.append(ids.numpy())
    if use_ids:
        parts.append(_u16(embed))
    raw = _run("norm", parts, tmp_path)
    y, res_out = _bf(raw[:M * d], (M, d)), _bf(raw[M * d:], (M, d))
    if use_ids:
        x = embed[ids]
    elif S > 0:
        acc = torch.zeros(M, d)
        for s in range(S):
            acc = acc + partial[s]
        x = acc.to(BF)
    else:
        x = dense
    if has_res:
        want_y, want_res = ops.rms_norm(x, w, eps, residual)
    else:
        want_y, want_res = ops.rms_norm(x, w, eps), x
    assert torch.equal(res_out, want_res)
    assert _ulp_frac(y, want_y) < 2e-3


@pytest.mark.parametrize("hd,H,KV,qk_norm,S", [(64, 4, 2, 0, 8), (128, 2, 1, 1, 3)])
def test_rope_store_source(tmp_path, hd, H, KV, qk_norm, S):
    g = torch.Generator().manual_seed(hd)
    M, max_pos, nslots, eps = 4, 64, 12, 1e-6
    table = ops.rope_table(hd, max_pos, 500000.0)
    pos = torch.tensor([5, 6, 33, 63], dtype=torch.int64)
    slots = torch.tensor([3, 9, -1, 0], dtype=torch.int32)  # one padded row: its K/V must not be written
    qn = (1 + 0.1 * torch.randn(hd, generator=g)).to(BF)
    kn = (1 + 0.1 * torch.randn(hd, generator=g)).to(BF)
    qkv_dim = (H + 2 * KV) * hd
    partial = torch.randn(S, M, qkv_dim, generator=g)
    raw = _run("rope", [np.array([M, H, KV, hd, S, max_pos, nslots, qk_norm], dtype=np.int32), np.array([eps], dtype=np.float32),
                        pos.numpy(), slots.numpy(), table.numpy().astype(np.float32), _u16(qn), _u16(kn),
                        partial.numpy().astype(np.float32)], tmp_path)
    nq, nc = M * H * hd, nslots * KV * hd
    q, kc, vc = _bf(raw[:nq], (M, H, hd)), _bf(raw[nq:nq + nc], (nslots, KV, hd)), _bf(raw[nq + nc:], (nslots, KV, hd))
    acc = torch.zeros(M, qkv_dim)
    for s in range(S):
        acc = acc + partial[s]
    qkv = acc.to(BF)
    wq, wk, wv = qkv.split([H * hd, KV * hd, KV * hd], -1)
    wq, wk, wv = wq.reshape(M, H, hd), wk.reshape(M, KV, hd), wv.reshape(M, KV, hd)
    if qk_norm:
        wq = ops.rms_norm(wq.reshape(-1, hd), qn, eps).reshape(M, H, hd)
        wk = ops.rms_norm(wk.reshape(-1, hd), kn, eps).reshape(M, KV, hd)
    wq, wk = ops.apply_rope(wq, pos, table), ops.apply_rope(wk, pos, table)
    assert _ulp_frac(q, wq) < 3e-3
    for m in range(M):
        if slots[m] >= 0:
            assert _ulp_frac(kc[slots[m]], wk[m]) < 1e-2
            assert torch.equal(vc[slots[m]], wv[m])
    written = {int(s) for s in slots if s >= 0}
    for s in range(nslots):
        if s not in written:
            assert not kc[s].any() and not vc[s].any(), "a slot that was not addressed was written"


def test_silu_mul_source(tmp_path):
    g = torch.Generator().manual_seed(3)
    M, ffn, S = 3, 512, 5
    partial = torch.randn(S, M, 2 * ffn, generator=g)
    raw = _run("silu", [np.array([M, ffn, S], dtype=np.int32), partial.numpy().astype(np.float32)], tmp_path)
    acc = torch.zeros(M, 2 * ffn)
    for s in range(S):
        acc = acc + partial[s]
    want = ops.silu_and_mul(acc.to(BF))
    assert _ulp_frac(_bf(raw, (M, ffn)), want) < 3e-3
