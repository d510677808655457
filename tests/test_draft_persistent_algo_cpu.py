"""CPU restatement of the ALGORITHM of csrc/draft_persistent.cuh (the experimental persistent draft forward), phase by
phase with the kernel's own index formulas — row dealing over (warp, CTA), the 16-byte lane/step layout of the GEMV,
rotate-half pairs, the (kv head, split) attention units with per-warp online softmax, the split merge, the residual
ping-pong and every bf16 rounding point — checked against the pinned oracle's decode forward.  This validates the design
(what is computed where, from which buffer, rounded when); the CUDA code itself still has to pass
tools/check_draft_persistent.py on a GPU before it leaves its opt-in switch."""
import numpy as np
import pytest
import torch

from oracle import ops
from oracle.model import ModelCfg, OracleModel, random_weights

BF = torch.bfloat16
N_CTAS, N_WARPS, N_SPLITS = 5, 8, 8  # gridDim.x, kDpWarps, kDpSplits


def r16(x):
    """round to bf16, keep as float32 (bf16_round in the kernel)"""
    return torch.as_tensor(np.asarray(x, dtype=np.float32)).to(BF).float().numpy()


def f32(t):
    return t.float().numpy()


def norm_prologue(a, b, w, eps):
    r = a.astype(np.float32) + (b.astype(np.float32) if b is not None else 0.0)
    resid_out = r16(r)
    rstd = np.float32(1.0) / np.sqrt(np.float32((r * r).sum(dtype=np.float32) / r.size + eps))
    return r16(r * rstd * w), resid_out


def gemv_rows(W, xs):
    """dp_gemv_rows / dp_dot2: row r belongs to warp r % (8 * grid) -> (warp, cta); lane l, step j covers elements
    [j * 256 + l * 8, +8); two accumulators per row (even / odd element), warp_sum, one bf16 rounding."""
    n, K = W.shape
    assert K % 256 == 0
    owner = np.full(n, -1)
    nw = N_WARPS * N_CTAS
    for warp in range(N_WARPS):
        for cta in range(N_CTAS):
            r0 = warp * N_CTAS + cta
            while r0 < n:
                for r in (r0, r0 + nw):
                    if r < n:
                        assert owner[r] == -1, "row dealt twice"
                        owner[r] = warp * N_CTAS + cta
                r0 += 2 * nw
    assert (owner >= 0).all(), "row never dealt"
    Wl = W.reshape(n, K // 256, 32, 8).astype(np.float32)
    xl = xs.reshape(K // 256, 32, 8)
    prod = Wl * xl[None]
    a0 = prod[..., 0::2].sum(axis=(1, 3), dtype=np.float32)  # [n, 32] per-lane partials, even elements
    a1 = prod[..., 1::2].sum(axis=(1, 3), dtype=np.float32)
    return r16((a0 + a1).sum(axis=1, dtype=np.float32))


def rope_row(x, cs, nw, eps, hd):
    half = hd // 2
    x1, x2 = x[:half].copy(), x[half:].copy()
    if nw is not None:
        rstd = np.float32(1.0) / np.sqrt(np.float32((x * x).sum(dtype=np.float32) / hd + eps))
        x1, x2 = r16(x1 * rstd * nw[:half]), r16(x2 * rstd * nw[half:])
    c, s = cs[:half], cs[half:]
    return np.concatenate([r16(x1 * c - x2 * s), r16(x2 * c + x1 * s)])


def attention_unit(cfg, qkv, h, s, ctx, pos, rope, kc, vc, bt, bs, lw, scale_log2):
    H, KV, hd = cfg.heads, cfg.kv_heads, cfg.head_dim
    G = H // KV
    cs = rope[pos]
    nq = f32(lw["q_norm"]) if cfg.qk_norm else None
    nk = f32(lw["k_norm"]) if cfg.qk_norm else None
    q = np.stack([rope_row(qkv[(h * G + g) * hd:(h * G + g + 1) * hd], cs, nq, cfg.rms_eps, hd) for g in range(G)])
    k_new = rope_row(qkv[(H + h) * hd:(H + h + 1) * hd], cs, nk, cfg.rms_eps, hd)
    v_new = qkv[(H + KV + h) * hd:(H + KV + h + 1) * hd].copy()
    store = None
    if s == 0:
        slot = bt[pos // bs] * bs + pos % bs
        store = (slot, k_new, v_new)
    per = -(-ctx // N_SPLITS)
    t0, t1 = s * per, min(ctx, s * per + per)
    m = np.full((N_WARPS, G), -np.inf, np.float32)
    l = np.zeros((N_WARPS, G), np.float32)
    acc = np.zeros((N_WARPS, G, hd), np.float32)
    for warp in range(N_WARPS):
        for t in range(t0 + warp, t1, N_WARPS):
            if t == pos:
                kv, vv = k_new, v_new
            else:
                slot = bt[t // bs] * bs + t % bs
                kv, vv = kc[slot, h], vc[slot, h]
            for g in range(G):
                sc = np.float32((q[g] * kv).sum(dtype=np.float32) * scale_log2)
                mn = max(m[warp, g], sc)
                corr = np.exp2(m[warp, g] - mn) if np.isfinite(m[warp, g]) else np.float32(0)
                pr = np.exp2(sc - mn)
                l[warp, g] = l[warp, g] * corr + pr
                acc[warp, g] = acc[warp, g] * corr + pr * vv
                m[warp, g] = mn
    out = np.zeros((G, hd + 2), np.float32)
    for g in range(G):
        mx = m[:, g].max()
        if np.isfinite(mx):
            wt = np.where(np.isfinite(m[:, g]), np.exp2(m[:, g] - mx), 0).astype(np.float32)
            out[g, :hd] = (acc[:, g] * wt[:, None]).sum(axis=0, dtype=np.float32)
            out[g, hd + 1] = (l[:, g] * wt).sum(dtype=np.float32)
        out[g, hd] = mx
    return out, store


def combine(part, H, hd):
    xs = np.zeros(H * hd, np.float32)
    for head in range(H):
        ms, ls, os_ = part[head, :, hd], part[head, :, hd + 1], part[head, :, :hd]
        mx = ms.max()
        wt = np.where(np.isfinite(ms), np.exp2(ms - mx), 0).astype(np.float32)
        o = (os_ * wt[:, None]).sum(axis=0, dtype=np.float32)
        lsum = (ls * wt).sum(dtype=np.float32)
        xs[head * hd:(head + 1) * hd] = r16(o / lsum) if lsum > 0 else 0
    return xs


def persistent_forward(cfg, w, kv_cache, rope, token, ctx0, pos_offset, bt, bs):
    """mirror of draft_forward_persistent_kernel; kv_cache [2, L, slots, KV, hd] float32 (updated in place)"""
    H, KV, hd, d = cfg.heads, cfg.kv_heads, cfg.head_dim, cfg.hidden
    ctx = ctx0 + pos_offset + 1
    pos = ctx - 1
    scale_log2 = np.float32(hd ** -0.5 * 1.4426950408889634)
    resid = [None, None]
    cur = 0
    vec_down = None
    emb = f32(w["embed"][token])
    for l, lw in enumerate(w["layers"]):
        if l == 0:
            xs, resid[cur ^ 1] = norm_prologue(emb, None, f32(lw["input_norm"]), cfg.rms_eps)
        else:
            xs, resid[cur ^ 1] = norm_prologue(vec_down, resid[cur], f32(lw["input_norm"]), cfg.rms_eps)
        cur ^= 1
        qkv = gemv_rows(f32(lw["qkv"]), xs)
        part = np.zeros((H, N_SPLITS, hd + 2), np.float32)
        G = H // KV
        stores = []
        for u in range(KV * N_SPLITS):
            h, s = u // N_SPLITS, u % N_SPLITS
            out, store = attention_unit(cfg, qkv, h, s, ctx, pos, rope, kv_cache[0, l], kv_cache[1, l], bt, bs, lw, scale_log2)
            part[h * G:(h + 1) * G, s] = out
            if store:
                stores.append((h, store))
        for h, (slot, k_new, v_new) in stores:  # units never read the new token from the cache, so order is free
            kv_cache[0, l, slot, h], kv_cache[1, l, slot, h] = k_new, v_new
        xs = combine(part, H, hd)
        vec_o = gemv_rows(f32(lw["o"]), xs)
        xs, resid[cur ^ 1] = norm_prologue(vec_o, resid[cur], f32(lw["post_norm"]), cfg.rms_eps)
        cur ^= 1
        gu = f32(lw["gate_up"])
        g, u = gemv_rows(gu[:cfg.ffn], xs), gemv_rows(gu[cfg.ffn:], xs)
        act = r16((g / (1.0 + np.exp(-g))) * u)
        vec_down = gemv_rows(f32(lw["down"]), act)
    xs, _ = norm_prologue(vec_down, resid[cur], f32(w["final_norm"]), cfg.rms_eps)
    return gemv_rows(f32(w["lm_head"]), xs)


@pytest.mark.parametrize("family", ["llama", "qwen"])
def test_persistent_draft_algorithm_matches_oracle_decode(family):
    torch.manual_seed(0)
    hd = 64 if family == "llama" else 128
    cfg = ModelCfg(hidden=256, layers=3, heads=4 if family == "llama" else 2, kv_heads=2 if family == "llama" else 1,
                   head_dim=hd, ffn=512, vocab=768, max_pos=512, rms_eps=1e-5 if family == "llama" else 1e-6,
                   rope_theta=500000.0, qk_norm=(family != "llama"))
    w = random_weights(cfg, seed=5)
    bs, nblk = 16, 8
    model = OracleModel(cfg, w, num_blocks=nblk, block_size=bs)
    bt = [3, 0, 5, 1, 7, 2, 6, 4]
    prompt = torch.randint(0, cfg.vocab, (37,))
    n = len(prompt)
    pos = torch.arange(n)
    slots = torch.tensor([bt[p // bs] * bs + p % bs for p in range(n)], dtype=torch.int32)
    btt = torch.tensor([bt], dtype=torch.int32)
    model.forward(prompt, pos, slots, torch.tensor([n], dtype=torch.int32), btt, n)  # fills the cache
    kv = model.kv_cache.float().numpy().reshape(2, cfg.layers, nblk * bs, cfg.kv_heads, hd).copy()
    rope = model.rope.numpy()
    tok = 123
    for step in range(3):  # three chained decode forwards, like the draft inside one speculative step
        p = n + step
        slot = torch.tensor([bt[p // bs] * bs + p % bs], dtype=torch.int32)
        hidden = model.forward(torch.tensor([tok]), torch.tensor([p]), slot, torch.tensor([p + 1], dtype=torch.int32), btt, 1)
        want = model.compute_logits(hidden)[0].float().numpy()
        got = persistent_forward(cfg, w, kv, rope, tok, n, step, bt, bs)
        scale = np.abs(want).max()
        assert np.abs(got - want).max() <= 0.02 * scale + 0.02, (step, np.abs(got - want).max(), scale)
        assert int(got.argmax()) == int(want.argmax())
        ref_kv = model.kv_cache.float().numpy().reshape(2, cfg.layers, nblk * bs, cfg.kv_heads, hd)
        assert np.abs(kv[:, :, int(slot)] - ref_kv[:, :, int(slot)]).max() <= 0.02 * np.abs(ref_kv[:, :, int(slot)]).max() + 1e-3
        tok = int(want.argmax())
