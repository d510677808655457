"""The SOURCE of csrc/sampling.cuh — the single-launch sampler and the single-launch verify() — compiled for the host
(tests/emu/cuda_emu.h) and held to the same bars as on the GPU (tests/test_ops_gpu.py): the reference's golden accept
counts / recovery tokens at temperature 0 (bit-exact, ties included), the oracle with the same Philox stream at
temperature > 0.  With SSD_B200_TSAN=1 the binary is built with ThreadSanitizer, which audits the kernels' ticket and
device-wide-barrier protocols for unsynchronised accesses."""
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

from tests.helpers import bf16, load

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "emu", "run_sampling.cpp")
TSAN = os.environ.get("SSD_B200_TSAN") == "1"
BIN = os.path.join(ROOT, "tests", "emu", "_build", "run_sampling" + ("_tsan" if TSAN else ""))

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++ (C++20)")


def _build():
    deps = [SRC, os.path.join(ROOT, "tests", "emu", "cuda_emu.h"), os.path.join(ROOT, "ssd_b200", "csrc", "sampling.cuh"),
            os.path.join(ROOT, "ssd_b200", "csrc", "common.cuh")]
    if os.path.exists(BIN) and all(os.path.getmtime(BIN) >= os.path.getmtime(d) for d in deps):
        return
    os.makedirs(os.path.dirname(BIN), exist_ok=True)
    flags = ["-fsanitize=thread", "-g"] if TSAN else []
    subprocess.run(["g++", "-std=c++20", "-O1", "-pthread", "-Wno-unknown-pragmas", "-Wno-attributes", *flags, "-o", BIN, SRC],
                   check=True)


def _u16(t):
    return t.contiguous().view(torch.int16).numpy().astype(np.uint16)


def _run(mode, blob_parts, tmp_path, out_counts):
    _build()
    inp, out = tmp_path / f"{mode}.in", tmp_path / f"{mode}.out"
    with open(inp, "wb") as f:
        for a in blob_parts:
            np.ascontiguousarray(a).tofile(f)
    res = subprocess.run([BIN, mode, str(inp), str(out)], capture_output=True, text=True, timeout=3000)
    assert res.returncode == 0, (res.returncode, res.stderr[-2000:])
    assert "ThreadSanitizer" not in res.stderr, res.stderr[:3000]
    raw = open(out, "rb").read()
    outs, off = [], 0
    for dtype, n in out_counts:
        sz = np.dtype(dtype).itemsize * n
        outs.append(np.frombuffer(raw[off:off + sz], dtype=dtype))
        off += sz
    return outs


def emu_sample(logits, temps, tmp_path, seed=0, step_id=0, nch=4):
    B, V = logits.shape
    (toks,) = _run("sample", [np.array([B, V, nch, seed, step_id], dtype=np.int64), temps.numpy().astype(np.float32),
                              _u16(logits)], tmp_path, [(np.int64, B)])
    return toks.tolist()


def emu_verify(lp, lq, spec, tt, tq, tmp_path, hits=None, jit=False, seed=0, step_id=0, nct=4):
    B, K1, V = lp.shape
    parts = [np.array([B, K1 - 1, V, nct, int(jit), int(hits is not None), seed, step_id], dtype=np.int64),
             tt.numpy().astype(np.float32), tq.numpy().astype(np.float32)]
    if hits is not None:
        parts.append(hits.numpy().astype(np.int32))
    parts += [spec.numpy().astype(np.int64), _u16(lp), _u16(lq)]
    n, rec = _run("verify", parts, tmp_path, [(np.int32, B), (np.int64, B)])
    return n.tolist(), rec.tolist()


def test_sampler_source_greedy_ties_and_temperature(tmp_path):
    from oracle import verify as V
    z = load("sampler_t0.npz")
    logits = bf16(z["logits"])
    assert emu_sample(logits, torch.zeros(logits.shape[0]), tmp_path) == z["tokens"].tolist()
    g = torch.Generator().manual_seed(5)
    logits = (torch.randn(4, 4096, generator=g) * 2).to(torch.bfloat16)
    temps = torch.tensor([0.7, 1.0, 0.0, 1.3])
    agree = 0
    for step in range(4):
        got = emu_sample(logits, temps, tmp_path, seed=123, step_id=step, nch=3 + step)
        want = V.sample(logits, temps, seed=123, call_id=step).tolist()
        agree += sum(int(a == b) for a, b in zip(got, want))
    assert agree >= 15  # same Philox stream; only fp32 log rounding can flip a near-tie


def test_verify_source_temp0_matches_reference_golden(tmp_path):
    z = load("verify_t0.npz")
    for c in range(int(z["n_cases"])):
        lp, lq, spec = bf16(z[f"c{c}_lp"]), bf16(z[f"c{c}_lq"]), torch.from_numpy(z[f"c{c}_spec"])
        B = lp.shape[0]
        n, rec = emu_verify(lp, lq, spec, torch.zeros(B), torch.zeros(B), tmp_path, nct=3 + c)
        assert n == z[f"c{c}_nacc"].tolist()
        assert rec == z[f"c{c}_rec"].tolist()


def test_verify_source_ratio_matches_oracle(tmp_path):
    from oracle import verify as V
    z = load("verify_ratio.npz")
    tot = ok = 0
    for c in range(int(z["n_cases"])):
        lp, lq, spec = bf16(z[f"c{c}_lp"]), bf16(z[f"c{c}_lq"]), torch.from_numpy(z[f"c{c}_spec"])
        tt, tq, jit = z[f"c{c}_cfg"].tolist()
        B = lp.shape[0]
        hits = torch.from_numpy(z[f"c{c}_hits"]) if f"c{c}_hits" in z else None
        for step in range(2):
            suf, rec = V.verify(lp, lq, spec, torch.full((B,), tt), torch.full((B,), tq), hits, bool(jit), None, 99, step)
            n, r = emu_verify(lp, lq, spec, torch.full((B,), tt), torch.full((B,), tq), tmp_path, hits, bool(jit), 99, step)
            tot += 2 * B
            ok += sum(int(a == len(s) - 1) for a, s in zip(n, suf))
            ok += sum(int(a == b) for a, b in zip(r, rec))
    assert ok >= tot - 2, f"{ok}/{tot}"
