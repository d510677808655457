"""Config — the keyword surface bench/bench.py:160-184 passes to LLM(...), same names and defaults as
ssd/config.py:7-49.  Options that belong to paths outside the sync-SD hot path (draft_async, use_eagle,
fan-out lists) are accepted and rejected loudly rather than silently ignored."""
from __future__ import annotations

import json
import os
from dataclasses import dataclass

from .paths import DEFAULT_DRAFT, DEFAULT_TARGET


@dataclass
class Config:
    model: str = DEFAULT_TARGET
    max_num_batched_tokens: int = 16384
    max_num_seqs: int = 1
    max_model_len: int = 4096
    gpu_memory_utilization: float = 0.7
    num_gpus: int = 1
    enforce_eager: bool = False
    hf_config: object | None = None
    eos: int = -1
    kvcache_block_size: int = 256
    num_kvcache_blocks: int = -1
    device: str = "cuda"
    # speculation
    draft_hf_config: object | None = None
    speculate: bool = False
    draft: str = DEFAULT_DRAFT
    speculate_k: int = 1
    draft_async: bool = False
    async_fan_out: int = 3
    fan_out_list: list[int] | None = None
    fan_out_list_miss: list[int] | None = None
    sampler_x: float | None = None
    jit_speculate: bool = False
    use_eagle: bool = False
    eagle_layers: list[int] | None = None
    d_model_target: int | None = None
    tokenizer_path: str | None = None
    verbose: bool = False
    debug_mode: bool = False
    max_steps: int | None = None
    # ssd_b200 extensions (ignored by the reference's bench scripts)
    use_cuda_graph: bool = True
    use_pdl: bool = True
    seed: int = 0

    @property
    def max_blocks(self) -> int:
        return (self.max_model_len + self.kvcache_block_size - 1) // self.kvcache_block_size

    def __post_init__(self):
        if not os.path.isdir(self.model):
            raise AssertionError(f"model directory {self.model!r} does not exist (config.py:53)")
        if not 1 <= self.num_gpus <= 8:
            raise AssertionError("single node only: 1 <= num_gpus <= 8 (config.py:55)")
        if self.draft_async:
            raise NotImplementedError("draft_async (async SSD) is a SURVEY §8(f) 'next' row, not built yet")
        if self.use_eagle:
            raise NotImplementedError("EAGLE-3 drafts are out of scope of the sync-SD hot path")
        if self.enforce_eager:
            self.use_cuda_graph = False
        self.hf_config = load_hf_config(self.model)
        self.max_model_len = min(self.max_model_len, self.hf_config.max_position_embeddings)
        if self.speculate:
            if not os.path.isdir(self.draft):
                raise AssertionError(f"draft directory {self.draft!r} does not exist")
            self.draft_hf_config = load_hf_config(self.draft)
            self.max_model_len = min(self.max_model_len, self.draft_hf_config.max_position_embeddings)
        if self.max_num_batched_tokens < self.max_model_len:
            raise AssertionError("max_num_batched_tokens < max_model_len (config.py:94)")
        # limits of libssdk (DESIGN.md "Limits"), reported here rather than at the first step
        k1 = (self.speculate_k + 1) if self.speculate else 1
        if self.speculate and not 1 <= self.speculate_k <= 7:
            raise ValueError(f"speculate_k={self.speculate_k}: libssdk supports 1 <= k <= 7")
        if self.max_num_seqs > 32 or self.max_num_seqs * k1 > 256:
            raise ValueError(f"max_num_seqs={self.max_num_seqs} with {k1} tokens per sequence and step: libssdk runs at most 32 "
                             "sequences and 256 tokens per step")


class _HFConfig:
    """Minimal attribute bag over config.json (what ssd/config.py reads through transformers.AutoConfig)."""

    def __init__(self, d: dict):
        self.__dict__.update(d)
        self.model_type = d.get("model_type", "llama")
        if "head_dim" not in d or d["head_dim"] is None:
            self.head_dim = d["hidden_size"] // d["num_attention_heads"]
        self.tie_word_embeddings = d.get("tie_word_embeddings", False)
        if "rope_theta" not in d:
            rp = d.get("rope_parameters") or {}
            self.rope_theta = rp.get("rope_theta", 1000000.0 if "qwen" in self.model_type else 500000.0)


def load_hf_config(path: str) -> _HFConfig:
    with open(os.path.join(path, "config.json")) as f:
        return _HFConfig(json.load(f))
