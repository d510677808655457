"""Sync speculative decoding driver over two OracleModels — restates SpeculatorSync.speculate
(engine/speculator_sync.py:25-69), Verifier.prefill/verify (engine/verifier.py:32-106), the
tensor prep of helpers/runner_helpers.py:50-121 and the sequence bookkeeping of
SpecDecodeStep.decode / Scheduler._update_sequence_metadata (engine/step.py:91-163,
engine/scheduler.py:248-262) for a batch of sequences with contiguous block tables."""
from __future__ import annotations

import torch

from . import verify as V
from .model import OracleModel


class SpecSession:
    def __init__(self, target: OracleModel, draft: OracleModel | None, K: int, max_blocks: int, seed: int = 0,
                 jit_speculate: bool = True):
        self.t, self.d, self.K, self.max_blocks, self.seed, self.jit = target, draft, K, max_blocks, seed, jit_speculate
        self.bs = target.block_size
        self.step_id = 0
        self.trace = []

    # -- helpers (runner_helpers.py:50-121) --
    def _prep(self, ctx0: list[int], q_len: int, block_tables: torch.Tensor):
        pos, slots = [], []
        for b, c0 in enumerate(ctx0):
            for j in range(q_len):
                p = c0 + j
                pos.append(p)
                slots.append(int(block_tables[b, p // self.bs]) * self.bs + p % self.bs)
        return (torch.tensor(pos, dtype=torch.int64), torch.tensor(slots, dtype=torch.int32),
                torch.tensor([c + q_len for c in ctx0], dtype=torch.int32))

    def _forward(self, model: OracleModel, ids: torch.Tensor, ctx0: list[int], q_len: int, bt: torch.Tensor):
        pos, slots, cl = self._prep(ctx0, q_len, bt)
        return model.forward(ids, pos, slots, cl, bt, q_len)

    def prefill(self, prompts: list[list[int]], temps: list[float], bt_t: torch.Tensor, bt_d: torch.Tensor | None):
        """verifier.py:32-52 + speculator_sync.py:14-23: one sequence at a time (prefill is off the hot path)."""
        self.ctx = [len(p) for p in prompts]
        self.bt_t, self.bt_d = bt_t, bt_d
        self.temps = temps
        rec = []
        for b, p in enumerate(prompts):
            ids = torch.tensor(p, dtype=torch.int64)
            h = self._forward(self.t, ids, [0], len(p), bt_t[b:b + 1])
            logits = self.t.compute_logits(h[-1:])
            tok = V.sample(logits, torch.tensor([temps[b]]), self.seed, self.step_id * 16 + 14)
            rec.append(int(tok[0]))
            if self.d is not None:
                self._forward(self.d, ids, [0], len(p), bt_d[b:b + 1])
        self.recovery = rec
        self.step_id += 1
        return rec

    def ar_step(self):
        """AutoRegressiveStep (engine/step.py:36-47) with the target only: feeds `recovery`, samples the next."""
        B = len(self.ctx)
        ids = torch.tensor(self.recovery, dtype=torch.int64)
        h = self._forward(self.t, ids, self.ctx, 1, self.bt_t)
        logits = self.t.compute_logits(h)
        nxt = V.sample(logits, torch.tensor(self.temps), self.seed, self.step_id * 16 + 14)
        out = list(self.recovery)
        self.ctx = [c + 1 for c in self.ctx]
        self.recovery = [int(t) for t in nxt]
        self.step_id += 1
        return out

    def spec_step(self, temps_q: list[float] | None = None):
        """One sync SD step. Returns (suffixes, recovery, logits_p, logits_q, speculations)."""
        B, K = len(self.ctx), self.K
        tq = torch.tensor(temps_q if temps_q is not None else self.temps, dtype=torch.float32)
        tt = torch.tensor(self.temps, dtype=torch.float32)
        spec = torch.zeros(B, K + 1, dtype=torch.int64)
        spec[:, 0] = torch.tensor(self.recovery)
        logits_q = []
        for k in range(K + 1):  # speculator_sync.py:47-65
            h = self._forward(self.d, spec[:, k].clone(), [c + k for c in self.ctx], 1, self.bt_d)
            if k == K:
                break  # last forward only writes KV (speculator_sync.py:55-56)
            lq = self.d.compute_logits(h)
            logits_q.append(lq)
            spec[:, k + 1] = V.sample(lq, tq, self.seed, self.step_id * 16 + k)
        logits_q = torch.stack(logits_q, dim=1)
        h = self._forward(self.t, spec.reshape(-1), self.ctx, K + 1, self.bt_t)  # verifier.py:65
        logits_p = self.t.compute_logits(h).view(B, K + 1, -1)
        suffixes, rec = V.verify(logits_p, logits_q, spec, tt, tq, None, self.jit, None, self.seed, self.step_id * 16 + 15)
        self.ctx = [c + len(s) for c, s in zip(self.ctx, suffixes)]  # scheduler.py:252-255
        self.recovery = rec
        self.step_id += 1
        return suffixes, rec, logits_p, logits_q, spec


    # ---- teacher-forced replay (tolerance protocol of SURVEY §7 "hard parts") ----
    def spec_step_forced(self, spec: torch.Tensor):
        """Run the K+1 draft forwards and the verify forward on GIVEN speculation tokens [B, K+1] without
        advancing the session. Returns (logits_p [B,K+1,V], logits_q [B,K,V])."""
        B, K = len(self.ctx), self.K
        logits_q = []
        for k in range(K + 1):
            h = self._forward(self.d, spec[:, k].clone(), [c + k for c in self.ctx], 1, self.bt_d)
            if k < K:
                logits_q.append(self.d.compute_logits(h))
        h = self._forward(self.t, spec.reshape(-1), self.ctx, K + 1, self.bt_t)
        return self.t.compute_logits(h).view(B, K + 1, -1), torch.stack(logits_q, dim=1)

    def advance(self, n_accept: list[int], recovery: list[int]):
        """Adopt an outcome (scheduler.py:252-262): ctx += accepted + 1, new recovery token."""
        self.ctx = [c + n + 1 for c, n in zip(self.ctx, n_accept)]
        self.recovery = list(recovery)
        self.step_id += 1


def top2_margin(logits: torch.Tensor) -> torch.Tensor:
    """fp32 gap between the best and second-best logit of each row (last dim)."""
    v = logits.float().topk(2, dim=-1).values
    return v[..., 0] - v[..., 1]


def check_greedy_step(spec: torch.Tensor, n_accept: list[int], recovery: list[int], lp: torch.Tensor, lq: torch.Tensor,
                      eps: float):
    """Compare a teacher's greedy (temp 0) step outcome with logits computed by a checker on the same tokens.
    Returns (hard_mismatches, near_tie_mismatches): a decision only counts as hard when the checker's own
    top-2 margin at that position is >= eps (bf16 logits tie within a few ulps far too often for exact match;
    SURVEY §7)."""
    B, Kp1 = spec.shape
    K = Kp1 - 1
    hard, soft = [], []
    mq, mp = top2_margin(lq), top2_margin(lp)
    aq, ap = lq.argmax(-1), lp.argmax(-1)

    def note(ok, margin, what):
        if ok:
            return
        (hard if float(margin) >= eps else soft).append((what, float(margin)))

    for b in range(B):
        for k in range(K):  # draft token k+1 was sampled from logits_q[:, k]
            note(int(spec[b, k + 1]) == int(aq[b, k]), mq[b, k], f"draft b{b} k{k}")
        n = n_accept[b]
        for j in range(n):  # accepted => target agreed
            note(int(spec[b, j + 1]) == int(ap[b, j]), mp[b, j], f"accept b{b} j{j}")
        if n < K:
            note(int(spec[b, n + 1]) != int(ap[b, n]), mp[b, n], f"reject b{b} j{n}")
        note(int(recovery[b]) == int(ap[b, n]), mp[b, n], f"recovery b{b}")
    return hard, soft


def contiguous_block_tables(B: int, max_blocks: int, offset: int = 0) -> torch.Tensor:
    """Sequence b owns blocks [offset + b*max_blocks, offset + (b+1)*max_blocks)."""
    return (torch.arange(B * max_blocks, dtype=torch.int32).view(B, max_blocks) + offset).contiguous()
