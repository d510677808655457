"""Scheduler — FCFS prefill / decode scheduling and spec-aware post-processing.

Behavioural contract of ssd/engine/scheduler.py:63-128 (schedule), :130-146 (preempt), :149-170 (postprocess),
:172-327 (postprocess_speculate): prefill has priority; decode reserves K+1 look-ahead slots in BOTH block
managers before a spec step; accepted suffixes are truncated at EOS / max_new_tokens / max_model_len; blocks the
suffix did not reach are returned; blocks that became full are hashed for prefix reuse."""
from __future__ import annotations

from collections import deque

from .block_manager import BlockManager
from .sequence import Sequence, SequenceStatus


class Scheduler:
    def __init__(self, config, draft_cfg=None):
        self.max_num_seqs = config.max_num_seqs
        self.max_num_batched_tokens = config.max_num_batched_tokens
        self.max_model_len = config.max_model_len
        self.eos = config.eos
        self.speculate = config.speculate
        self.K = config.speculate_k
        self.block_size = config.kvcache_block_size
        self.block_manager = BlockManager(config.num_kvcache_blocks, self.block_size, is_draft=False,
                                          max_model_len=self.max_model_len)
        self.draft_block_manager = None
        if self.speculate:
            nb = getattr(draft_cfg, "num_kvcache_blocks", config.num_kvcache_blocks)
            self.draft_block_manager = BlockManager(nb, self.block_size, is_draft=True, speculate_k=self.K,
                                                    max_model_len=self.max_model_len)
        self.waiting: deque[Sequence] = deque()
        self.running: deque[Sequence] = deque()
        self.retired: list[Sequence] = []  # finished inside schedule() (no room for another step); drained by LLMEngine.step

    def _managers(self):
        return [m for m in (self.block_manager, self.draft_block_manager) if m is not None]

    def is_finished(self) -> bool:
        return not self.waiting and not self.running

    def add(self, seq: Sequence) -> None:
        self.waiting.append(seq)

    # ---------------------------------------------------------------------------------- schedule
    def schedule(self) -> tuple[list[Sequence], bool]:
        admitted: list[Sequence] = []
        budget = self.max_num_batched_tokens
        while self.waiting:
            seq = self.waiting[0]
            need = len(seq) - seq.num_cached_tokens
            if need > budget or not all(m.can_allocate(seq) for m in self._managers()):
                break
            for m in self._managers():
                m.allocate(seq)
            budget -= need
            seq.status = SequenceStatus.RUNNING
            self.running.append(self.waiting.popleft())
            admitted.append(seq)
        if admitted:
            return admitted, True

        lookahead = self.K + 1 if self.speculate else 1
        chosen: list[Sequence] = []
        while self.running and len(chosen) < self.max_num_seqs:
            seq = self.running.popleft()
            if seq.num_tokens + lookahead > self.max_model_len:
                # No room for another (speculative) step below max_model_len.  The reference's can_append() answers
                # False for this case forever: schedule() preempts, the sequence is re-prefilled and preempted again
                # (scheduler.py:101-118, block_manager.py:150-152) and generate() spins.  Finish the sequence instead.
                self._retire(seq)
                continue
            ok = True
            while not all(m.can_append(seq, lookahead) for m in self._managers()):
                if self.running:
                    self.preempt(self.running.pop())
                else:
                    self.preempt(seq)
                    ok = False
                    break
            if ok:
                for m in self._managers():
                    m.may_append(seq, lookahead)
                chosen.append(seq)
        self.running.extendleft(reversed(chosen))
        return chosen, False

    def preempt(self, seq: Sequence) -> None:
        """Out of KV blocks: drop the sequence's cache and re-prefill it later (completions become prompt)."""
        seq.status = SequenceStatus.WAITING
        seq.recovery_token_id = None
        for m in self._managers():
            m.deallocate(seq)
        seq.num_prompt_tokens = seq.num_tokens
        seq.last_spec_step_accepted_len = -1
        self.waiting.appendleft(seq)

    # ------------------------------------------------------------------------- autoregressive
    def postprocess(self, seqs: list[Sequence], token_ids: list[int], is_prefill: bool) -> None:
        for seq, tok in zip(seqs, token_ids):
            seq.append_token(tok)
            seq.num_cached_tokens = seq.num_prompt_tokens if is_prefill else seq.num_cached_tokens + 1
            done = (not seq.ignore_eos and tok == self.eos) or seq.num_completion_tokens == seq.max_new_tokens
            if done:
                self._finish(seq)
            elif seq.num_tokens % self.block_size == 0:
                self.block_manager.seal(seq, seq.num_blocks - 1)

    # ------------------------------------------------------------------------------ speculative
    def _truncate(self, seq: Sequence, suffix: list[int]) -> tuple[list[int], bool]:
        if not seq.ignore_eos and self.eos in suffix:
            suffix = suffix[:suffix.index(self.eos) + 1]
        room = seq.max_new_tokens - seq.num_completion_tokens
        if len(suffix) >= room:
            suffix = suffix[:room]
        if seq.num_tokens + len(suffix) > self.max_model_len:
            suffix = suffix[:max(0, self.max_model_len - seq.num_tokens)]
        finished = ((not seq.ignore_eos and self.eos in suffix)
                    or seq.num_completion_tokens + len(suffix) == seq.max_new_tokens
                    or seq.num_tokens + len(suffix) >= self.max_model_len)
        return suffix, finished

    def postprocess_speculate(self, seqs: list[Sequence], new_suffixes: list[list[int]], next_recovery_tokens: list[int],
                              eagle_acts=None) -> None:
        for seq, suffix, recovery in zip(seqs, new_suffixes, next_recovery_tokens):
            suffix, finished = self._truncate(seq, suffix)
            if not suffix:
                raise AssertionError("accepted suffix is empty")
            old_full = seq.num_tokens // self.block_size
            keep = -(-(seq.num_tokens + len(suffix)) // self.block_size)
            for m in self._managers():
                m.trim(seq, keep)
            seq.token_ids.extend(suffix)
            seq.num_cached_tokens += len(suffix)
            seq.num_draft_cached_tokens += len(suffix)
            seq.last_spec_step_accepted_len = len(suffix)
            seq.recovery_token_id = recovery
            for idx in range(old_full, seq.num_tokens // self.block_size):
                for m in self._managers():
                    if m.blocks[m._table(seq)[idx]].digest == -1:
                        m.seal(seq, idx)
            if finished:
                self._finish(seq)

    def _retire(self, seq: Sequence) -> None:
        seq.status = SequenceStatus.FINISHED
        for m in self._managers():
            m.deallocate(seq)
        self.retired.append(seq)

    def _finish(self, seq: Sequence) -> None:
        seq.status = SequenceStatus.FINISHED
        for m in self._managers():
            m.deallocate(seq)
        self.running.remove(seq)
