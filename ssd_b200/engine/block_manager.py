"""BlockManager — paged-KV bookkeeping on the host with hash-chained prefix reuse.

Same contract as ssd/engine/block_manager.py:26-177 (allocate / deallocate / can_append / may_append, xxh64 of
a full block chained on the previous block's hash, reference counts, FIFO free list) because the block tables
it produces are hot-path inputs; written from scratch around a small PagePool."""
from __future__ import annotations

from collections import deque

import numpy as np
import xxhash


class _Page:
    __slots__ = ("pid", "refs", "digest", "tokens")

    def __init__(self, pid: int):
        self.pid, self.refs, self.digest, self.tokens = pid, 0, -1, None


class BlockManager:
    def __init__(self, num_blocks: int, block_size: int, is_draft: bool = False, speculate_k: int = -1,
                 max_model_len: int = -1, verbose: bool = False):
        if num_blocks <= 0:
            raise ValueError("KV cache has no blocks")
        self.block_size, self.is_draft, self.max_model_len = block_size, is_draft, max_model_len
        self.blocks = [_Page(i) for i in range(num_blocks)]
        self.free_block_ids: deque[int] = deque(range(num_blocks))
        self.used_block_ids: set[int] = set()
        self.hash_to_block_id: dict[int, int] = {}

    # -- hashing -------------------------------------------------------------------------------
    @staticmethod
    def compute_hash(token_ids: list[int], prefix: int = -1) -> int:
        h = xxhash.xxh64()
        if prefix != -1:
            h.update(prefix.to_bytes(8, "little"))
        h.update(np.array(token_ids).tobytes())
        return h.intdigest()

    # -- page pool -----------------------------------------------------------------------------
    def _table(self, seq):
        return seq.draft_block_table if self.is_draft else seq.block_table

    def _take(self, pid: int) -> _Page:
        page = self.blocks[pid]
        assert page.refs == 0
        page.refs, page.digest, page.tokens = 1, -1, None
        self.free_block_ids.remove(pid)
        self.used_block_ids.add(pid)
        return page

    def _release(self, pid: int) -> None:
        page = self.blocks[pid]
        page.refs -= 1
        if page.refs == 0:
            self.used_block_ids.discard(pid)
            self.free_block_ids.append(pid)

    def seal(self, seq, index: int) -> None:
        """A block of `seq` became full: record its chained hash so later prompts can reuse it."""
        table = self._table(seq)
        tokens = seq.block(index)
        prev = self.blocks[table[index - 1]].digest if index > 0 else -1
        page = self.blocks[table[index]]
        page.digest = self.compute_hash(tokens, prev)
        page.tokens = tokens
        self.hash_to_block_id[page.digest] = page.pid

    # -- sequence-level API ----------------------------------------------------------------------
    def can_allocate(self, seq) -> bool:
        return len(self.free_block_ids) >= seq.num_blocks

    def allocate(self, seq) -> None:
        table = self._table(seq)
        assert not table
        digest, missed = -1, False
        for i in range(seq.num_blocks):
            tokens = seq.block(i)
            full = len(tokens) == self.block_size
            digest = self.compute_hash(tokens, digest) if full else -1
            pid = self.hash_to_block_id.get(digest, -1)
            if pid == -1 or self.blocks[pid].tokens != tokens:
                missed = True
            if missed:
                page = self._take(self.free_block_ids[0])
            else:
                if self.is_draft:
                    seq.num_draft_cached_tokens += self.block_size
                else:
                    seq.num_cached_tokens += self.block_size
                page = self.blocks[pid]
                if pid in self.used_block_ids:
                    page.refs += 1
                else:
                    page = self._take(pid)
            if digest != -1:
                page.digest, page.tokens = digest, tokens
                self.hash_to_block_id[digest] = page.pid
            table.append(page.pid)

    def deallocate(self, seq) -> None:
        table = self._table(seq)
        for pid in reversed(table):
            self._release(pid)
        table.clear()
        if self.is_draft:
            seq.num_draft_cached_tokens = 0
        else:
            seq.num_cached_tokens = 0

    def _blocks_for(self, seq, lookahead: int) -> int:
        return -(-(seq.num_tokens + lookahead) // self.block_size)

    def can_append(self, seq, lookahead_num_tokens: int = 1) -> bool:
        if seq.num_tokens + lookahead_num_tokens > self.max_model_len:
            return False
        extra = self._blocks_for(seq, lookahead_num_tokens) - len(self._table(seq))
        return extra <= 0 or len(self.free_block_ids) >= extra

    def may_append(self, seq, lookahead_num_tokens: int = 1) -> None:
        table = self._table(seq)
        extra = self._blocks_for(seq, lookahead_num_tokens) - len(table)
        if extra > len(self.free_block_ids):
            raise RuntimeError(f"out of KV blocks: need {extra}, have {len(self.free_block_ids)}")
        for _ in range(max(0, extra)):
            table.append(self._take(self.free_block_ids[0]).pid)

    def trim(self, seq, keep_blocks: int) -> None:
        """Give back look-ahead blocks that the accepted suffix did not reach (scheduler.py:205-240)."""
        table = self._table(seq)
        while len(table) > keep_blocks:
            self._release(table.pop())
