"""Sequence — per-request host state.  Field names follow ssd/engine/sequence.py:14-120 because the scheduler,
the block managers and bench scripts address them by name; the implementation is independent."""
from __future__ import annotations

import enum
import itertools

from ..sampling_params import SamplingParams


class SequenceStatus(enum.Enum):
    WAITING = 1
    RUNNING = 2
    FINISHED = 3


_ids = itertools.count()


class Sequence:
    block_size = 256  # set by LLMEngine from Config.kvcache_block_size (llm_engine.py:46)

    def __init__(self, token_ids: list[int], sampling_params: SamplingParams | None = None):
        sp = sampling_params or SamplingParams()
        self.seq_id = next(_ids)
        self.status = SequenceStatus.WAITING
        self.token_ids = list(token_ids)
        self.num_prompt_tokens = len(self.token_ids)
        # KV bookkeeping: tokens whose K/V are valid in the target / draft cache
        self.num_cached_tokens = 0
        self.num_draft_cached_tokens = 0
        self.block_table: list[int] = []
        self.draft_block_table: list[int] = []
        self.temperature = sp.temperature
        self.draft_temperature = sp.draft_temperature
        self.max_new_tokens = sp.max_new_tokens
        self.ignore_eos = sp.ignore_eos
        # speculation state: the token sampled by the last verify/prefill that is not yet in token_ids
        self.recovery_token_id: int | None = None
        self.last_spec_step_accepted_len = -1

    # -- views --------------------------------------------------------------------------------
    def __len__(self) -> int:
        return len(self.token_ids)

    def __getitem__(self, key):
        return self.token_ids[key]

    @property
    def num_tokens(self) -> int:
        return len(self.token_ids)

    @property
    def last_token(self) -> int:
        return self.token_ids[-1]

    @property
    def is_finished(self) -> bool:
        return self.status is SequenceStatus.FINISHED

    @property
    def num_completion_tokens(self) -> int:
        return len(self.token_ids) - self.num_prompt_tokens

    @property
    def prompt_token_ids(self) -> list[int]:
        return self.token_ids[:self.num_prompt_tokens]

    @property
    def completion_token_ids(self) -> list[int]:
        return self.token_ids[self.num_prompt_tokens:]

    @property
    def num_blocks(self) -> int:
        return -(-len(self.token_ids) // self.block_size)

    def block(self, i: int) -> list[int]:
        bs = self.block_size
        return self.token_ids[i * bs:(i + 1) * bs]

    def append_token(self, token_id: int) -> None:
        self.token_ids.append(token_id)

    @property
    def effective_draft_temperature(self) -> float:
        """verifier.py:85-88: draft temperature falls back to the target temperature."""
        return self.temperature if self.draft_temperature is None else self.draft_temperature
