"""Result containers of the speculate / verify plug-in points (ssd/engine/helpers/speculate_types.py:7-46)."""
from __future__ import annotations

from abc import ABC, abstractmethod
from dataclasses import dataclass

import numpy as np


@dataclass
class SpeculateResult:
    speculations: np.ndarray  # [B, K+1] int64: column 0 is the recovery token
    logits_q: object | None   # stays on the device inside libssdk (ssdk_logits_q); None on the host side
    cache_hits: object | None = None


@dataclass
class VerifyResult:
    new_suffixes: list[list[int]]
    recovery_tokens: list[int]
    eagle_acts: object | None = None


class SpeculatorBase(ABC):
    def __init__(self, lookahead: int, device):
        self.lookahead, self.device = lookahead, device

    @abstractmethod
    def prefill(self, seqs, verify_result): ...

    @abstractmethod
    def speculate(self, seqs, verify_result): ...


class VerifierBase(ABC):
    def __init__(self, lookahead: int, device):
        self.lookahead, self.device = lookahead, device

    @abstractmethod
    def prefill(self, seqs, eagle: bool = False): ...

    @abstractmethod
    def verify(self, seqs, speculate_result, eagle: bool = False): ...
