"""Restatement of the reference's decoder forward (models/llama3.py:89-99,130-134,185-199,248-273,
models/qwen3.py:90-108) over a paged KV cache, on CPU in bf16.  One class covers both families:
Qwen3 = Llama + per-head q/k RMSNorm before RoPE.  Weights are held in the reference's packed
per-rank layout (what utils/loader.py + the weight_loader callbacks of layers/linear.py produce)."""
from __future__ import annotations

from dataclasses import dataclass

import torch

from . import ops

BF16 = torch.bfloat16


@dataclass
class ModelCfg:
    hidden: int
    layers: int
    heads: int
    kv_heads: int
    head_dim: int
    ffn: int
    vocab: int
    rms_eps: float = 1e-5
    rope_theta: float = 500000.0
    qk_norm: bool = False
    tie_embed: bool = False
    max_pos: int = 8192


def random_weights(cfg: ModelCfg, seed: int, std: float = 0.02, norm_jitter: float = 0.1) -> dict:
    """Seeded synthetic weights in packed layout. Norm weights are 1 + jitter so that weight-ordering bugs show."""
    g = torch.Generator().manual_seed(seed)

    def mat(r, c, s=std):
        return (torch.randn(r, c, generator=g) * s).to(BF16)

    def vec(n):
        return (1.0 + norm_jitter * torch.randn(n, generator=g)).to(BF16)

    w = {"embed": mat(cfg.vocab, cfg.hidden, 1.0), "final_norm": vec(cfg.hidden), "layers": []}
    w["lm_head"] = w["embed"] if cfg.tie_embed else mat(cfg.vocab, cfg.hidden, 0.05)
    qkv_dim = (cfg.heads + 2 * cfg.kv_heads) * cfg.head_dim
    for _ in range(cfg.layers):
        lw = {
            "input_norm": vec(cfg.hidden),
            "qkv": mat(qkv_dim, cfg.hidden, 0.05),
            "o": mat(cfg.hidden, cfg.heads * cfg.head_dim, 0.05),
            "post_norm": vec(cfg.hidden),
            "gate_up": mat(2 * cfg.ffn, cfg.hidden, 0.05),
            "down": mat(cfg.hidden, cfg.ffn, 0.05),
        }
        if cfg.qk_norm:
            lw["q_norm"] = vec(cfg.head_dim)
            lw["k_norm"] = vec(cfg.head_dim)
        w["layers"].append(lw)
    return w


class OracleModel:
    def __init__(self, cfg: ModelCfg, weights: dict, num_blocks: int, block_size: int = 256, compiled: bool = True):
        self.cfg, self.w, self.block_size, self.compiled = cfg, weights, block_size, compiled
        self.rope = ops.rope_table(cfg.head_dim, cfg.max_pos, cfg.rope_theta)
        # ModelRunner.allocate_kv_cache (engine/model_runner.py:484-491): [2, L, nblk, bs, KV, hd]
        self.kv_cache = torch.zeros(2, cfg.layers, num_blocks, block_size, cfg.kv_heads, cfg.head_dim, dtype=BF16)

    def forward(self, input_ids, positions, slot_mapping, context_lens, block_tables, q_len: int) -> torch.Tensor:
        """LlamaModel.forward (models/llama3.py:248-273); attention always reads K/V back from the paged cache
        (layers/attention.py:82-83,107-111)."""
        c, w = self.cfg, self.w
        H, KV, hd = c.heads, c.kv_heads, c.head_dim
        hidden = w["embed"][input_ids]  # VocabParallelEmbedding tp=1 (embed_head.py:49-58)
        residual = None
        for l, lw in enumerate(w["layers"]):
            if residual is None:  # llama3.py:192-193
                hidden, residual = ops.rms_norm(hidden, lw["input_norm"], c.rms_eps, compiled=self.compiled), hidden
            else:
                hidden, residual = ops.rms_norm(hidden, lw["input_norm"], c.rms_eps, residual, compiled=self.compiled)
            qkv = ops.linear(hidden, lw["qkv"])  # llama3.py:94
            q, k, v = qkv.split([H * hd, KV * hd, KV * hd], dim=-1)
            N = q.shape[0]
            q, k, v = q.reshape(N, H, hd), k.reshape(N, KV, hd), v.reshape(N, KV, hd)
            if c.qk_norm:  # qwen3.py:97-103
                q = ops.rms_norm(q.reshape(-1, hd), lw["q_norm"], c.rms_eps, compiled=self.compiled).reshape(N, H, hd)
                k = ops.rms_norm(k.reshape(-1, hd), lw["k_norm"], c.rms_eps, compiled=self.compiled).reshape(N, KV, hd)
            q = ops.apply_rope(q, positions, self.rope)
            k = ops.apply_rope(k, positions, self.rope)
            ops.store_kvcache(k, v, self.kv_cache[0, l], self.kv_cache[1, l], slot_mapping)
            o = ops.paged_attention(q, self.kv_cache[0, l], self.kv_cache[1, l], block_tables, context_lens, q_len,
                                    hd ** -0.5)
            attn = ops.linear(o, lw["o"])  # llama3.py:98
            hidden, residual = ops.rms_norm(attn, lw["post_norm"], c.rms_eps, residual, compiled=self.compiled)
            gu = ops.linear(hidden, lw["gate_up"])  # llama3.py:131
            hidden = ops.linear(ops.silu_and_mul(gu, compiled=self.compiled), lw["down"])  # :132-133
        hidden, _ = ops.rms_norm(hidden, w["final_norm"], c.rms_eps, residual, compiled=self.compiled)  # :266
        return hidden

    def compute_logits(self, hidden: torch.Tensor) -> torch.Tensor:
        """ParallelLMHead.forward tp=1 (embed_head.py:94-116): bf16 logits."""
        return ops.linear(hidden, self.w["lm_head"])
