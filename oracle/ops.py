"""Layer-level restatements (torch on CPU, bf16 storage) of ssd/layers/*.

`compiled=True` (default) follows the numerics the reference actually runs with — its
RMSNorm / SiluAndMul are @torch.compile regions, and Inductor keeps the in-kernel
`.to(bf16)` in fp32, i.e. ONE rounding (SURVEY §8a checklist 2); `compiled=False` is the eager
(double-rounding) reading of the same source, used to pin against eager goldens.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

BF16 = torch.bfloat16


def linear(x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """layers/linear.py:98,196 / embed_head.py:95,111 — F.linear on bf16: fp32 accumulate, one bf16 rounding."""
    return F.linear(x.to(BF16), w.to(BF16))


def rms_norm(x: torch.Tensor, w: torch.Tensor, eps: float, residual: torch.Tensor | None = None, compiled: bool = True):
    """RMSDNorm / RMSHeadNorm (layers/layernorm.py:16-40,64-88).

    residual given: r = x + residual (fp32), new_residual = bf16(r) (:83-84); statistics use the unrounded r."""
    r = x.float()
    new_res = None
    if residual is not None:
        r = r + residual.float()
        new_res = r.to(BF16)
    var = r.pow(2).mean(dim=-1, keepdim=True)
    n = r * torch.rsqrt(var + eps)
    if compiled:
        y = (n * w.float()).to(BF16)
    else:
        y = n.to(BF16) * w.to(BF16)
    return y if residual is None else (y, new_res)


def rope_table(head_dim: int, max_pos: int, base: float) -> torch.Tensor:
    """RotaryEmbedding.__init__ (layers/rotary_embedding.py:30-37): fp32 [max_pos, hd] = cos | sin."""
    inv_freq = 1.0 / (base ** (torch.arange(0, head_dim, 2, dtype=torch.float) / head_dim))
    t = torch.arange(max_pos, dtype=torch.float)
    freqs = torch.einsum("i,j -> ij", t, inv_freq)
    return torch.cat((freqs.cos(), freqs.sin()), dim=-1)


def apply_rope(x: torch.Tensor, positions: torch.Tensor, table: torch.Tensor) -> torch.Tensor:
    """apply_rotary_emb (layers/rotary_embedding.py:6-17, 47-59): x [N, heads, hd] bf16, NeoX halves, fp32 math."""
    cs = table[positions]
    cos, sin = cs.chunk(2, dim=-1)
    cos, sin = cos.unsqueeze(-2), sin.unsqueeze(-2)
    x1, x2 = torch.chunk(x.float(), 2, dim=-1)
    y1 = x1 * cos - x2 * sin
    y2 = x2 * cos + x1 * sin
    return torch.cat((y1, y2), dim=-1).to(x.dtype)


def silu_and_mul(x: torch.Tensor, compiled: bool = True) -> torch.Tensor:
    """SiluAndMul (layers/activation.py:11-14)."""
    g, u = x.chunk(2, -1)
    if compiled:
        return (F.silu(g.float()) * u.float()).to(BF16)
    return F.silu(g) * u


def store_kvcache(k: torch.Tensor, v: torch.Tensor, k_cache: torch.Tensor, v_cache: torch.Tensor, slot_mapping: torch.Tensor):
    """store_kvcache (layers/attention.py:10-41): caches [nblk, bs, KV, hd] viewed as [slots, KV, hd]; slot -1 skipped."""
    KV, hd = k.shape[-2], k.shape[-1]
    kc, vc = k_cache.view(-1, KV, hd), v_cache.view(-1, KV, hd)
    keep = slot_mapping >= 0
    idx = slot_mapping[keep].long()
    kc[idx] = k[keep]
    vc[idx] = v[keep]


def paged_attention(q: torch.Tensor, k_cache: torch.Tensor, v_cache: torch.Tensor, block_tables: torch.Tensor,
                    context_lens: torch.Tensor, q_len: int, scale: float) -> torch.Tensor:
    """flash_attn_with_kvcache as called at layers/attention.py:107-111 (verify, cu_seqlens_q step q_len) and
    :128-131 (decode): causal aligned to the END of cache_seqlens, GQA head h -> kv head h // (H/KV),
    pages looked up through block_tables.  fp32 softmax.  (The arithmetic lives in sgl-kernel 0.3.17.post1 /
    FlashAttention-3, which is not vendored in the reference; this restates its published semantics.)

    q [B*q_len, H, hd] -> [B*q_len, H*hd]."""
    Mq, H, hd = q.shape
    B = Mq // q_len
    _, bs, KV, _ = k_cache.shape
    G = H // KV
    out = torch.empty(Mq, H, hd, dtype=q.dtype)
    for b in range(B):
        L = int(context_lens[b])
        nb = (L + bs - 1) // bs
        pages = block_tables[b, :nb].long()
        k = k_cache[pages].reshape(nb * bs, KV, hd)[:L].float()
        v = v_cache[pages].reshape(nb * bs, KV, hd)[:L].float()
        k = k.repeat_interleave(G, dim=1)  # [L, H, hd]
        v = v.repeat_interleave(G, dim=1)
        qb = q[b * q_len:(b + 1) * q_len].float()  # [q, H, hd]
        s = torch.einsum("qhd,lhd->hql", qb, k) * scale
        qpos = torch.arange(q_len).unsqueeze(1) + (L - q_len)
        mask = torch.arange(L).unsqueeze(0) <= qpos  # [q, L]
        s = s.masked_fill(~mask.unsqueeze(0), float("-inf"))
        p = torch.softmax(s, dim=-1)
        out[b * q_len:(b + 1) * q_len] = torch.einsum("hql,lhd->qhd", p, v).to(q.dtype)
    return out.reshape(Mq, H * hd)
