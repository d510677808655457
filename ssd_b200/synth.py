"""Synthetic model directories: exact Llama-3 / Qwen-3 shapes, seeded random weights, no checkpoint download.

There are no model weights in the image, so throughput is measured on synthetic weights at the true shapes
(every kernel streams the real byte count).  To make speculative decoding behave like a real target/draft pair the
weights are built with the "bigram agreement" construction of SURVEY §8(d):

  * token embeddings e_t ~ N(0,1)^d (near-orthogonal), lm_head[pi(t)] = e_t  => greedy next token = pi(t);
  * o_proj / down_proj are scaled to ~0 so the residual stream stays the embedding (they are still streamed);
  * the draft knows the target's next token for a seeded fraction `alpha` of tokens (and guesses noise otherwise), so the accepted
    prefix length is Geometric(alpha) truncated at K:  E[tokens/step] = (1 - alpha^(K+1)) / (1 - alpha).

A directory holds config.json (HF layout), ssd_b200_synthetic.json (seed, alpha, role) and a WordLevel
tokenizer.json; the weights are generated ON THE DEVICE by `generate_weights` at load time.
"""
from __future__ import annotations

import json
import os

import torch

SHAPES = {
    # name: hidden, layers, heads, kv_heads, head_dim, ffn, vocab, rms_eps, rope_theta, model_type, tied
    "llama-3.2-1b": (2048, 16, 32, 8, 64, 8192, 128256, 1e-5, 500000.0, "llama", True),
    "llama-3.1-8b": (4096, 32, 32, 8, 128, 14336, 128256, 1e-5, 500000.0, "llama", False),
    "llama-3.1-70b": (8192, 80, 64, 8, 128, 28672, 128256, 1e-5, 500000.0, "llama", False),
    "qwen3-0.6b": (1024, 28, 16, 8, 128, 3072, 151936, 1e-6, 1000000.0, "qwen3", True),
    "qwen3-32b": (5120, 64, 64, 8, 128, 25600, 151936, 1e-6, 1000000.0, "qwen3", False),
    # tiny shapes for smoke tests
    "llama-tiny-target": (256, 2, 4, 2, 64, 512, 1024, 1e-5, 500000.0, "llama", False),
    "llama-tiny-draft": (128, 1, 2, 1, 64, 256, 1024, 1e-5, 500000.0, "llama", False),
}


def make_model_dir(root: str, shape: str, role: str, seed: int = 0, alpha: float = 0.85, layers: int | None = None,
                   max_position_embeddings: int = 131072, o_down_std: float = 1e-5, rng: str = "torch",
                   draft_mode: str = "perm", lm_scale: float | None = None) -> str:
    """Write <root>/<family>-synthetic-<shape>-<role>/ and return its path.  `role` is "target" or "draft"."""
    h, L, H, KV, hd, ffn, V, eps, theta, mtype, tied = SHAPES[shape]
    L = layers or L
    family = "llama" if mtype == "llama" else "qwen"
    path = os.path.join(root, f"{family}-synthetic-{shape}-{role}")
    os.makedirs(path, exist_ok=True)
    cfg = {
        "model_type": mtype, "architectures": ["LlamaForCausalLM" if mtype == "llama" else "Qwen3ForCausalLM"],
        "hidden_size": h, "num_hidden_layers": L, "num_attention_heads": H, "num_key_value_heads": KV, "head_dim": hd,
        "intermediate_size": ffn, "vocab_size": V, "rms_norm_eps": eps, "rope_theta": theta, "rope_scaling": None,
        "max_position_embeddings": max_position_embeddings, "hidden_act": "silu", "torch_dtype": "bfloat16",
        # real Llama-3.2-1B / Qwen3-0.6B tie embeddings; the synthetic pair unties them so the draft can carry its own
        # next-token permutation — per-forward HBM traffic is unchanged (embedding lookups read M rows either way)
        "tie_word_embeddings": False, "real_checkpoint_ties_embeddings": tied,
        "bos_token_id": 0, "eos_token_id": 1,
    }
    with open(os.path.join(path, "config.json"), "w") as f:
        json.dump(cfg, f, indent=1)
    with open(os.path.join(path, "ssd_b200_synthetic.json"), "w") as f:
        json.dump({"seed": seed, "alpha": alpha, "role": role, "shape": shape, "o_down_std": o_down_std,
                   "rng": rng, "draft_mode": draft_mode, "lm_scale": lm_scale}, f)
    return path


class SyntheticTokenizer:
    """Token i <-> the string "<i>"; enough for LLMEngine (encode/decode/eos_token_id)."""

    def __init__(self, path: str):
        with open(os.path.join(path, "config.json")) as f:
            cfg = json.load(f)
        self.vocab_size = cfg["vocab_size"]
        self.eos_token_id = cfg.get("eos_token_id", 1)

    def encode(self, text: str) -> list[int]:
        out = []
        for part in text.replace(">", "> ").split():
            part = part.strip("<>")
            out.append(int(part) % self.vocab_size if part.isdigit() else (hash(part) % self.vocab_size))
        return out or [0]

    def decode(self, ids) -> str:
        return "".join(f"<{int(i)}>" for i in ids)


def permutations(vocab: int, seed: int, alpha: float, device, draft_mode: str = "perm") -> tuple[torch.Tensor, torch.Tensor]:
    """(pi_target, pi_draft): pi_draft[t] == pi_target[t] on a seeded fraction alpha of tokens.

    draft_mode "perm" (default): on the other tokens the draft follows a DIFFERENT next-token map with the same large
    margin — pi_draft(t) = pi_target(sigma(t)) with sigma a cyclic shift inside the disagreeing set — so pi_draft is a
    permutation too, every draft decision is as clear-cut as the target's and a whole run is reproducible token for
    token (engine vs oracle vs the reference engine on the same weights).  "noise" (round 1): pi_draft = -1 there, the
    draft's lm_head row is missing and its guess is the argmax of noise (a near-tie a few per cent of the time)."""
    g = torch.Generator(device="cpu").manual_seed(seed * 7919 + 13)
    pi_t = torch.randperm(vocab, generator=g, device="cpu")  # explicit: the caller may run under set_default_device("cuda")
    agree = torch.rand(vocab, generator=g, device="cpu") < alpha
    if draft_mode == "noise":
        pi_d = torch.where(agree, pi_t, torch.full_like(pi_t, -1))
    else:
        pi_d = pi_t.clone()
        dis = (~agree).nonzero().squeeze(1)
        if dis.numel() > 1:
            pi_d[dis] = pi_t[torch.roll(dis, -1)]
    return pi_t.to(device), pi_d.to(device)


def hash_uniform(rows: int, cols: int, std: float, stream: int, device, row0: int = 0) -> torch.Tensor:
    """Device-independent pseudo-random bf16 matrix [rows, cols], zero mean, standard deviation `std` (uniform): element
    (r, c) is a pure integer function of (stream, r + row0, c) — splitmix64-style, wrapping int64 arithmetic — so the
    same bits come out on CPU and GPU.  Used where a fixture generated in the (GPU-less) build container must be
    regenerated bit-for-bit on the GPU box (tests/golden/true_width_*.npz)."""
    out = torch.empty(rows, cols, dtype=torch.bfloat16, device=device)
    step = max(1, (1 << 25) // max(cols, 1))
    c = torch.arange(cols, dtype=torch.int64, device=device)[None, :]
    scale = std * (3.0 ** 0.5) / 32768.0
    off = (stream * 0x51ED270B27B4F3) & 0x3FFFFFFFFFFFFFFF  # keep the Python int inside int64
    for r in range(0, rows, step):
        n = min(step, rows - r)
        idx = (torch.arange(r + row0, r + row0 + n, dtype=torch.int64, device=device)[:, None] * 1000003 + c
               + off)
        x = idx * -7046029254386353131  # 0x9E3779B97F4A7C15 as int64
        x = x ^ ((x >> 30) & 0x3FFFFFFFF)
        x = x * -4658895280553007687    # 0xBF58476D1CE4E5B9
        x = x ^ ((x >> 27) & 0x1FFFFFFFFF)
        x = x * -7723592293110705685    # 0x94D049BB133111EB
        x = x ^ ((x >> 31) & 0x1FFFFFFFF)
        u = ((x >> 20) & 0xFFFF) - 32768  # [-32768, 32767]
        out[r:r + n] = (u.to(torch.float32) * scale).to(torch.bfloat16)
    return out


def iter_weights(spec, meta: dict, device, tp_size: int = 1, tp_rank: int = 0):
    """Yield the packed per-rank bf16 weights of a synthetic directory one piece at a time, in RNG order:
    ("embed", t), ("lm_head", t), ("final_norm", t), then ("layer", l, {...}) for every decoder layer.  A consumer that
    copies each piece into its own storage and drops it never holds two copies of a 70B model (the reference GPU arm
    fills the reference's nn.Parameters this way); `generate_weights` simply collects the pieces.

    Sharding follows the reference's rules (layers/linear.py:90-95,116-122,148-162,188-193; embed_head.py:41-47):
    column-parallel qkv / gate_up by output rows per head group, row-parallel o / down by input columns,
    embedding and lm_head by vocab rows."""
    seed, alpha, role = meta["seed"], meta["alpha"], meta["role"]
    od_std = float(meta.get("o_down_std", 1e-5))
    o_std, down_std = float(meta.get("o_std", od_std)), float(meta.get("down_std", od_std))
    hashed = meta.get("rng", "torch") == "hash"
    role_id = 1 if role == "target" else 2
    d, V = spec.hidden, spec.vocab
    H, KV, hd, ffn = spec.heads // tp_size, spec.kv_heads // tp_size, spec.head_dim, spec.ffn // tp_size
    Vs = V // tp_size
    bf = torch.bfloat16
    g = torch.Generator(device=device).manual_seed(seed * 1000003 + 17)  # embeddings: same for target and draft

    def randn(rows, cols, std, gen):
        out = torch.empty(rows, cols, dtype=bf, device=device)
        step = max(1, (1 << 27) // max(cols, 1))  # generate in <=128M-element slabs to bound fp32 temporaries
        for r in range(0, rows, step):
            n = min(step, rows - r)
            out[r:r + n] = (torch.randn(n, cols, generator=gen, device=device, dtype=torch.float32) * std).to(bf)
        return out

    if hashed:
        n_mat = [0]

        def randn(rows, cols, std, gen):  # noqa: F811 — same call sites, device-independent values
            n_mat[0] += 1
            return hash_uniform(rows, cols, std, seed * 4096 + (0 if gen is g else role_id * 1024 + tp_rank * 128) + n_mat[0],
                                device)

    embed_full = randn(V, d, 1.0, g)  # identical on every rank and for both roles
    pi_t, pi_d = permutations(V, seed, alpha, device, meta.get("draft_mode", "perm"))
    pi = pi_t if role == "target" else pi_d
    lo, hi = tp_rank * Vs, (tp_rank + 1) * Vs
    # lm_head[pi(t)] = e_t  <=>  lm_head[v] = e_{pi^-1(v)}.  Draft: only the agreeing tokens get their row; the rows
    # of the others stay zero, so for a disagreeing token the draft's argmax is a noise token != pi_t(t).
    if role == "target":
        inv = torch.empty_like(pi)
        inv[pi] = torch.arange(V, device=device)
        lm_shard = embed_full[inv[lo:hi]].contiguous()
    else:
        lm_head = torch.zeros(V, d, dtype=bf, device=device)
        known = (pi >= 0).nonzero().squeeze(1)
        lm_head[pi[known]] = embed_full[known]
        lm_shard = lm_head[lo:hi].contiguous()
        del lm_head
    lm_scale = meta.get("lm_scale")
    if lm_scale:
        # soften the next-token distribution for temp > 0 runs: logit of the bigram successor ~= lm_scale, the others
        # ~ N(0, lm_scale^2 / d) (unscaled: ~d against ~sqrt(d), i.e. a one-hot softmax at any sane temperature)
        for r0 in range(0, lm_shard.shape[0], 16384):
            lm_shard[r0:r0 + 16384] = (lm_shard[r0:r0 + 16384].float() * (float(lm_scale) / d)).to(bf)
    embed_shard = embed_full[lo:hi].contiguous()
    del embed_full
    yield ("embed", embed_shard)
    del embed_shard
    yield ("lm_head", lm_shard)
    del lm_shard
    yield ("final_norm", torch.ones(d, dtype=bf, device=device))
    gl = torch.Generator(device=device).manual_seed(seed * 1000003 + (101 if role == "target" else 202) + 7 * tp_rank)
    for l in range(spec.layers):
        lw = {
            "input_norm": torch.ones(d, dtype=bf, device=device),
            "post_norm": torch.ones(d, dtype=bf, device=device),
            "qkv": randn((H + 2 * KV) * hd, d, 0.02, gl),
            # default ~0: keeps the residual stream == embedding (accept-len control); `o_down_std` in the
            # directory's meta switches the attention / MLP branches on (parity tests)
            "o": randn(d, H * hd, o_std, gl),
            "gate_up": randn(2 * ffn, d, 0.02, gl),
            "down": randn(d, ffn, down_std, gl),
        }
        if spec.qk_norm:
            lw["q_norm"] = torch.ones(hd, dtype=bf, device=device)
            lw["k_norm"] = torch.ones(hd, dtype=bf, device=device)
        yield ("layer", l, lw)


def generate_weights(spec, meta: dict, device, tp_size: int = 1, tp_rank: int = 0) -> dict:
    """Packed per-rank bf16 weights on `device` for a synthetic directory (`meta` = ssd_b200_synthetic.json)."""
    w = {"layers": []}
    for item in iter_weights(spec, meta, device, tp_size, tp_rank):
        if item[0] == "layer":
            w["layers"].append(item[2])
        else:
            w[item[0]] = item[1]
    return w


def expected_tokens_per_step(alpha: float, K: int) -> float:
    return (1 - alpha ** (K + 1)) / (1 - alpha) if alpha < 1 else K + 1
