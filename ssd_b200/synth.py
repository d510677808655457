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
                   max_position_embeddings: int = 131072) -> str:
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
        json.dump({"seed": seed, "alpha": alpha, "role": role, "shape": shape}, f)
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


def permutations(vocab: int, seed: int, alpha: float, device) -> tuple[torch.Tensor, torch.Tensor]:
    """(pi_target, pi_draft): pi_draft[t] == pi_target[t] on a seeded fraction alpha of tokens."""
    g = torch.Generator(device="cpu").manual_seed(seed * 7919 + 13)
    pi_t = torch.randperm(vocab, generator=g)
    agree = torch.rand(vocab, generator=g) < alpha
    pi_d = torch.where(agree, pi_t, torch.full_like(pi_t, -1))  # -1: the draft has no idea (its guess will be noise)
    return pi_t.to(device), pi_d.to(device)


def generate_weights(spec, meta: dict, device, tp_size: int = 1, tp_rank: int = 0) -> dict:
    """Packed per-rank bf16 weights on `device` for a synthetic directory (`meta` = ssd_b200_synthetic.json).

    Sharding follows the reference's rules (layers/linear.py:90-95,116-122,148-162,188-193; embed_head.py:41-47):
    column-parallel qkv / gate_up by output rows per head group, row-parallel o / down by input columns,
    embedding and lm_head by vocab rows."""
    seed, alpha, role = meta["seed"], meta["alpha"], meta["role"]
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

    embed_full = randn(V, d, 1.0, g)  # identical on every rank and for both roles
    pi_t, pi_d = permutations(V, seed, alpha, device)
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
    w = {"embed": embed_full[lo:hi].contiguous(), "lm_head": lm_shard,
         "final_norm": torch.ones(d, dtype=bf, device=device), "layers": []}
    del embed_full
    gl = torch.Generator(device=device).manual_seed(seed * 1000003 + (101 if role == "target" else 202) + 7 * tp_rank)
    for _ in range(spec.layers):
        lw = {
            "input_norm": torch.ones(d, dtype=bf, device=device),
            "post_norm": torch.ones(d, dtype=bf, device=device),
            "qkv": randn((H + 2 * KV) * hd, d, 0.02, gl),
            "o": randn(d, H * hd, 1e-5, gl),        # ~0: keeps the residual stream == embedding
            "gate_up": randn(2 * ffn, d, 0.02, gl),
            "down": randn(d, ffn, 1e-5, gl),
        }
        if spec.qk_norm:
            lw["q_norm"] = torch.ones(hd, dtype=bf, device=device)
            lw["k_norm"] = torch.ones(hd, dtype=bf, device=device)
        w["layers"].append(lw)
    return w


def expected_tokens_per_step(alpha: float, K: int) -> float:
    return (1 - alpha ** (K + 1)) / (1 - alpha) if alpha < 1 else K + 1
