"""Restatement of ssd.utils.verify.verify (utils/verify.py:5-181) and ssd.layers.sampler.Sampler
(layers/sampler.py:14-36) with the random draws made explicit.

The decision logic (greedy prefix, ratio rows, acceptance probabilities, residual distribution,
which recovery token is kept) follows the reference line by line; the three places where the
reference consumes torch's global RNG are parameterised:
  * acceptance uniforms  (verify.py:115  torch.rand_like)      -> `uniforms` [B,K] or Philox(TAG_ACCEPT)
  * recovery draws       (verify.py:158-159 torch.multinomial) -> exponential race with Philox(TAG_RECOVER)
  * sampler exponentials (sampler.py:33 exponential_)          -> Philox(TAG_SAMPLE)
torch.multinomial(p, 1) is itself argmax(p / Exp(1)), so the race is the same distribution.
"""
from __future__ import annotations

import numpy as np
import torch

from . import philox


def _softmax_rows(logits: torch.Tensor, temps: torch.Tensor) -> torch.Tensor:
    """verify.py:74-86 / :89-100 — fp32 softmax(logits / T) for T>0 rows, one-hot at the bf16 argmax for T==0 rows."""
    B = logits.shape[0]
    probs = torch.zeros(logits.shape, dtype=torch.float32)
    for b in range(B):
        t = float(temps[b])
        if t > 0:
            tt = torch.tensor(max(t, 1e-8), dtype=torch.float32)
            probs[b] = torch.softmax(logits[b].float() / tt, dim=-1)  # bf16 / dimensioned fp32 tensor promotes to fp32
        else:
            am = logits[b].argmax(dim=-1)
            probs[b].scatter_(1, am.unsqueeze(-1), 1.0)
    return probs


def verify(logits_p: torch.Tensor, logits_q: torch.Tensor, speculations: torch.Tensor, temperatures_target: torch.Tensor,
           temperatures_draft: torch.Tensor, cache_hits: torch.Tensor | None = None, jit_speculate: bool = False,
           uniforms: torch.Tensor | None = None, seed: int = 0, call_id: int = 0, return_debug: bool = False):
    """Returns (accepted_suffixes: list[list[int]], recovery_tokens: list[int]) like the reference.

    call_id is the device kernel's Philox call id (ssdk_verify: step_id; engine: step*16+15)."""
    B, Kp1, V = logits_p.shape
    K = Kp1 - 1
    draft_tokens = speculations[:, 1:]
    # 1) greedy path (verify.py:29-48)
    preds_p = logits_p.argmax(dim=-1)
    matches = draft_tokens == preds_p[:, :-1]
    accept_greedy = torch.full((B,), K, dtype=torch.int64)
    for b in range(B):
        mism = (~matches[b]).nonzero()
        if mism.numel():
            accept_greedy[b] = int(mism[0])
    rec_greedy = preds_p[torch.arange(B), accept_greedy]

    temps_t = temperatures_target.float()
    temps_q = temperatures_draft.float()
    base_ratio_rows = (temps_t > 0) | (temps_q > 0)  # verify.py:57
    if jit_speculate:
        ratio_rows = base_ratio_rows
    else:
        hits = cache_hits.to(torch.bool) if cache_hits is not None else torch.zeros(B, dtype=torch.bool)
        ratio_rows = base_ratio_rows & hits
    do_any_ratio = bool(ratio_rows.any())
    need_p = bool((temps_t > 0).any()) or do_any_ratio

    dbg = {}
    probs_p = _softmax_rows(logits_p, temps_t) if need_p else None
    accept_until = accept_greedy.clone()
    probs_q = None
    if do_any_ratio:
        probs_q = _softmax_rows(logits_q, temps_q)
        idx = draft_tokens.unsqueeze(2)
        p_vals = probs_p[:, :K, :].gather(2, idx).squeeze(2)
        q_vals = probs_q.gather(2, idx).squeeze(2)
        accept_probs = (p_vals / (q_vals + 1e-10)).clamp(max=1.0)  # verify.py:114
        if uniforms is None:
            js = np.arange(K, dtype=np.uint64)[None, :].repeat(B, 0)
            rows = np.arange(B, dtype=np.uint64)[:, None].repeat(K, 1)
            w = philox.draw(js, rows, call_id, philox.TAG_ACCEPT, seed)[0]
            uniforms = torch.from_numpy(philox.unit_half_open(w))
        accepts = uniforms <= accept_probs  # verify.py:116
        accept_ratio = torch.full((B,), K, dtype=torch.int64)
        for b in range(B):
            rej = (~accepts[b]).nonzero()
            if rej.numel():
                accept_ratio[b] = int(rej[0])
        accept_until = torch.where(ratio_rows, accept_ratio, accept_greedy)  # verify.py:127
        dbg["accept_probs"] = accept_probs
        dbg["uniforms"] = uniforms

    # 3) recovery distribution (verify.py:137-164)
    rec_final = rec_greedy.clone()
    dists = [None] * B
    if probs_p is not None:
        for b in range(B):
            if not temps_t[b] > 0:
                continue  # verify.py:167: temp==0 rows keep the greedy recovery
            n = int(accept_until[b])
            p_fallback = probs_p[b, n]
            fallback = p_fallback / p_fallback.sum()
            dist = fallback
            stream = 1  # Philox word used by the device for the fallback race
            if do_any_ratio and bool(ratio_rows[b]) and n < K:
                q_slice = probs_q[b, min(n, K - 1)]
                adj = (p_fallback - q_slice).clamp(min=0.0)
                s = adj.sum()
                if s > 0:
                    dist = adj / s
                    stream = 0
            dists[b] = dist
            words = philox.draw(np.arange(V, dtype=np.uint64), np.uint64(b), call_id, philox.TAG_RECOVER, seed)
            e = torch.from_numpy(philox.exp1(words[stream]))
            score = torch.where(dist > 0, torch.log(dist) - torch.log(e), torch.full_like(dist, float("-inf")))
            rec_final[b] = int(score.argmax())
    dbg["recovery_dists"] = dists
    dbg["accept_until"] = accept_until

    suffixes = []
    for b in range(B):
        n = int(accept_until[b])
        suffixes.append([int(speculations[b, 0])] + [int(t) for t in draft_tokens[b, :n]])
    out = (suffixes, [int(t) for t in rec_final])
    return out + (dbg,) if return_debug else out


def sample(logits: torch.Tensor, temperatures: torch.Tensor, seed: int = 0, call_id: int = 0) -> torch.Tensor:
    """Sampler.forward (layers/sampler.py:14-36): greedy where T == 0, else argmax(softmax(l/T) / (Exp(1)+1e-10)),
    evaluated in the log domain with the device's Philox exponentials (TAG_SAMPLE, word = index % 4)."""
    B, V = logits.shape
    l = logits.float()
    out = l.argmax(dim=-1)
    for b in range(B):
        t = float(temperatures[b])
        if t == 0:
            continue
        i = np.arange(V, dtype=np.uint64)
        words = philox.draw(i >> np.uint64(2), np.uint64(b), call_id, philox.TAG_SAMPLE, seed)
        w = np.stack(words, axis=-1)[np.arange(V), (i & np.uint64(3)).astype(np.int64)]
        e = torch.from_numpy(philox.exp1(w)) + 1e-10
        inv_t = torch.tensor(1.0 / t, dtype=torch.float32)
        score = l[b] * inv_t - torch.log(e)
        out[b] = int(score.argmax())
    return out
