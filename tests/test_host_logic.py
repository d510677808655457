"""CPU tests of the host-side mirror of the reference's scheduler / block manager / sequence bookkeeping
(the producers of the hot path's inputs: block tables, ctx lens, recovery tokens)."""
import random
import types

import pytest

from ssd_b200.engine.block_manager import BlockManager
from ssd_b200.engine.scheduler import Scheduler
from ssd_b200.engine.sequence import Sequence, SequenceStatus
from ssd_b200.sampling_params import SamplingParams


def _cfg(**kw):
    d = dict(max_num_seqs=2, max_num_batched_tokens=4096, max_model_len=1024, eos=1, speculate=True, speculate_k=4,
             kvcache_block_size=16, num_kvcache_blocks=64)
    d.update(kw)
    return types.SimpleNamespace(**d)


@pytest.fixture(autouse=True)
def _bs():
    Sequence.block_size = 16
    yield
    Sequence.block_size = 256


def test_block_hash_is_xxh64_chain():
    import numpy as np
    import xxhash
    toks = list(range(16))
    h0 = BlockManager.compute_hash(toks)
    assert h0 == xxhash.xxh64(np.array(toks).tobytes()).intdigest()
    h = xxhash.xxh64()
    h.update(h0.to_bytes(8, "little"))
    h.update(np.array(toks).tobytes())
    assert BlockManager.compute_hash(toks, h0) == h.intdigest()


def test_allocate_prefix_reuse_and_refcounts():
    bm = BlockManager(8, 16, max_model_len=1024)
    a = Sequence(list(range(40)))
    bm.allocate(a)
    assert len(a.block_table) == 3 and a.num_cached_tokens == 0
    b = Sequence(list(range(32)) + [99] * 5)  # shares two full blocks with a
    bm.allocate(b)
    assert b.block_table[:2] == a.block_table[:2] and b.block_table[2] != a.block_table[2]
    assert b.num_cached_tokens == 32
    assert bm.blocks[a.block_table[0]].refs == 2
    bm.deallocate(a)
    assert bm.blocks[b.block_table[0]].refs == 1
    bm.deallocate(b)
    assert len(bm.free_block_ids) == 8 and not bm.used_block_ids


def test_may_append_reserves_lookahead_and_trim_returns_it():
    bm = BlockManager(8, 16, max_model_len=1024)
    s = Sequence(list(range(30)))
    bm.allocate(s)
    assert len(s.block_table) == 2
    assert bm.can_append(s, 5)
    bm.may_append(s, 5)  # 30 + 5 = 35 tokens -> 3 blocks
    assert len(s.block_table) == 3
    bm.trim(s, 2)
    assert len(s.block_table) == 2 and len(bm.free_block_ids) == 6
    assert not BlockManager(8, 16, max_model_len=32).can_append(s, 5)


def test_spec_postprocess_truncation_rules():
    sch = Scheduler(_cfg(), draft_cfg=types.SimpleNamespace(num_kvcache_blocks=64))
    s = Sequence(list(range(10, 20)), SamplingParams(temperature=0.0, max_new_tokens=6, ignore_eos=False))
    sch.add(s)
    seqs, is_prefill = sch.schedule()
    assert is_prefill and seqs == [s] and s.block_table and s.draft_block_table
    s.recovery_token_id = 7
    s.num_cached_tokens = s.num_draft_cached_tokens = 10
    seqs, is_prefill = sch.schedule()
    assert not is_prefill
    sch.postprocess_speculate([s], [[7, 8, 9]], [5])
    assert s.token_ids[-3:] == [7, 8, 9] and s.recovery_token_id == 5 and s.num_cached_tokens == 13
    # EOS inside the suffix: truncate after it and finish
    sch.schedule()
    sch.postprocess_speculate([s], [[5, 1, 4, 4]], [3])
    assert s.token_ids[-2:] == [5, 1] and s.is_finished and not s.block_table and not sch.running


def test_spec_postprocess_max_new_tokens_and_block_sealing():
    sch = Scheduler(_cfg(), draft_cfg=types.SimpleNamespace(num_kvcache_blocks=64))
    s = Sequence(list(range(100, 114)), SamplingParams(temperature=0.0, max_new_tokens=7, ignore_eos=True))
    sch.add(s)
    sch.schedule()
    s.recovery_token_id = 2
    s.num_cached_tokens = s.num_draft_cached_tokens = 14
    sch.schedule()
    assert len(s.block_table) == 2  # 14 + K + 1 = 19 tokens -> look-ahead block reserved
    sch.postprocess_speculate([s], [[2, 3, 4]], [9])  # crosses the 16-token boundary: block 0 is sealed
    assert sch.block_manager.blocks[s.block_table[0]].digest != -1
    assert sch.draft_block_manager.blocks[s.draft_block_table[0]].digest != -1
    sch.schedule()
    sch.postprocess_speculate([s], [[9, 9, 9, 9, 9]], [0])  # only 4 tokens of room left
    assert s.num_completion_tokens == 7 and s.is_finished


def test_scheduler_random_walk_invariants():
    """Random accept lengths: block tables always cover ctx + K + 1 slots at step time, nothing leaks."""
    rng = random.Random(0)
    sch = Scheduler(_cfg(num_kvcache_blocks=48, max_num_seqs=3), draft_cfg=types.SimpleNamespace(num_kvcache_blocks=48))
    seqs = [Sequence([rng.randrange(2, 50) for _ in range(rng.randrange(5, 40))],
                     SamplingParams(temperature=0.0, max_new_tokens=60, ignore_eos=True)) for _ in range(5)]
    for s in seqs:
        sch.add(s)
    steps = 0
    while not sch.is_finished() and steps < 500:
        steps += 1
        batch, is_prefill = sch.schedule()
        assert batch
        if is_prefill:
            for s in batch:
                s.recovery_token_id = 3
                s.num_cached_tokens = s.num_draft_cached_tokens = s.num_prompt_tokens
            continue
        for s in batch:
            need = -(-(s.num_tokens + 5) // 16)
            assert len(s.block_table) >= need and len(s.draft_block_table) >= need
            assert s.num_cached_tokens == s.num_tokens
        sch.postprocess_speculate(batch, [[s.recovery_token_id] + [4] * rng.randrange(0, 5) for s in batch], [3] * len(batch))
    assert sch.is_finished()
    assert all(s.num_completion_tokens == 60 for s in seqs)
    for bm in (sch.block_manager, sch.draft_block_manager):
        assert len(bm.free_block_ids) == 48 and not bm.used_block_ids


def test_no_room_below_max_model_len_finishes_instead_of_spinning():
    """A sequence whose length lands in (max_model_len - K - 1, max_model_len) cannot take another spec step.  The
    reference preempts and re-prefills it forever (scheduler.py:101-118); here schedule() retires it."""
    sch = Scheduler(_cfg(num_kvcache_blocks=48, max_num_seqs=2, max_model_len=64),
                    draft_cfg=types.SimpleNamespace(num_kvcache_blocks=48))
    s = Sequence(list(range(2, 42)), SamplingParams(temperature=0.0, max_new_tokens=500, ignore_eos=True))
    sch.add(s)
    batch, is_prefill = sch.schedule()
    assert is_prefill and batch == [s]
    s.recovery_token_id = 3
    s.num_cached_tokens = s.num_draft_cached_tokens = s.num_prompt_tokens
    steps = 0
    while not sch.is_finished() and steps < 50:
        steps += 1
        batch, is_prefill = sch.schedule()
        assert not is_prefill
        if not batch:
            break
        sch.postprocess_speculate(batch, [[s.recovery_token_id, 4, 4]], [3])  # 3 tokens per step: 40 -> 58 -> stop
    assert sch.is_finished() and sch.retired == [s] and s.is_finished
    assert 64 - 5 <= s.num_tokens <= 64
    for bm in (sch.block_manager, sch.draft_block_manager):
        assert len(bm.free_block_ids) == 48 and not bm.used_block_ids


def test_compat_shim_exposes_reference_names():
    import ssd_b200.compat as cm
    cm.install()
    from ssd import LLM, SamplingParams as SP  # noqa: F401
    from ssd.engine.llm_engine import METRICS
    import ssd.paths as P
    for k in ("cache_hits", "accepted_suffix_lens_with_recovery", "prefill_total_time", "decode_total_time",
              "prefill_total_tokens", "decode_total_tokens", "target_step_times", "target_verify_times"):
        assert k in METRICS
    assert hasattr(P, "DATASET_PATHS") and hasattr(P, "HF_CACHE_DIR") and hasattr(P, "EAGLE3_QWEN_32B")
    assert SP().temperature == 1.0 and SP().max_new_tokens == 256 and SP().draft_temperature is None


def test_prefill_call_planner_packs_equal_prompts_and_never_loses_tokens():
    """PairRunner.plan_prefill_call / prefill_many (host side of the varlen prefill, runner_helpers.py:123-180): every
    prompt token is run exactly once and in order, a call never exceeds the token or batch limit, equal-length prompts
    share calls, and prefix-cache hits are kept out of the shared calls (their pages may still be in flight)."""
    from ssd_b200.runner import PairRunner

    class Rec:
        max_batch = 8
        plan_prefill_call = staticmethod(PairRunner.plan_prefill_call)
        prefill_many = PairRunner.prefill_many
        prefill = PairRunner.prefill

        def __init__(self):
            self.calls = []

        def forward_tokens(self, which, ids, ctx_len, block_tables, temps=None, want_sample=True, seed=0):
            assert len(ids) <= self.max_batch and len(ids) * len(ids[0]) <= 256 and len({len(x) for x in ids}) == 1
            self.calls.append((ids, list(ctx_len), [bt[0] for bt in block_tables], want_sample))
            return [x[-1] + 1000 for x in ids] if want_sample else None

    assert PairRunner.plan_prefill_call([128] * 16, 256, 32) == ([0, 1], 128)
    assert PairRunner.plan_prefill_call([300, 40, 30], 256, 32) == ([0], 256)
    assert PairRunner.plan_prefill_call([44, 40, 30], 256, 32) == ([0, 1, 2], 30)
    assert PairRunner.plan_prefill_call([0, 0], 256, 32) == ([], 0)
    assert PairRunner.plan_prefill_call([5] * 40, 256, 8)[0] == list(range(8))

    rng = random.Random(3)
    for trial in range(20):
        n = rng.randint(1, 12)
        lens = [rng.choice([1, 7, 128, 128, 300, 517]) for _ in range(n)]
        starts = [rng.choice([0, 0, 0, min(256, l - 1)]) for l in lens]
        toks = [[i * 10000 + j for j in range(l)] for i, l in enumerate(lens)]
        r = Rec()
        out = r.prefill_many(0, toks, [[i] for i in range(n)], starts, want_sample=True)
        assert out == [t[-1] + 1000 for t in toks]
        seen = {i: starts[i] for i in range(n)}
        for ids, ctx, owners, _ in r.calls:
            if len(ids) > 1:
                assert all(starts[o] == 0 for o in owners), "a prefix-cache hit shared a call"
            for row, c, o in zip(ids, ctx, owners):
                assert c == seen[o] and row == toks[o][c:c + len(row)]
                seen[o] += len(row)
        assert all(seen[i] == lens[i] for i in range(n))
    r = Rec()
    r.prefill_many(0, [list(range(128))] * 16, [[i] for i in range(16)], [0] * 16)
    assert len(r.calls) == 8


def test_kv_budget_draft_cache_is_sized_from_what_the_target_left(monkeypatch):
    """loader.kv_blocks_for (allocate_kv_cache, engine/model_runner.py:446-476; draft_runner.py:27): both caches are carved
    out of ONE free-memory snapshot, so the draft's share is taken from what remains after the target's; the result never
    exceeds what max_num_seqs sequences of max_model_len can use and never drops below one block."""
    import torch
    from ssd_b200 import loader
    spec_t = types.SimpleNamespace(layers=80, kv_heads=8, head_dim=128)
    spec_d = types.SimpleNamespace(layers=16, kv_heads=8, head_dim=64)
    cfg = types.SimpleNamespace(kvcache_block_size=256, gpu_memory_utilization=0.9, max_num_seqs=64, max_blocks=32)
    free = 30 << 30
    monkeypatch.setattr(torch.cuda, "mem_get_info", lambda *a: (free, 180 << 30))
    bt = loader.kv_block_bytes(cfg, spec_t, 1)
    assert bt == 2 * 80 * 256 * 8 * 128 * 2 and loader.kv_block_bytes(cfg, spec_t, 4) == bt // 4
    nbt = loader.kv_blocks_for(cfg, spec_t, 1, 0.8)
    assert nbt == int(free * 0.9 * 0.8) // bt
    nbd = loader.kv_blocks_for(cfg, spec_d, 1, 0.75, reserved=nbt * bt)
    bd = loader.kv_block_bytes(cfg, spec_d, 1)
    assert nbd == int((free - nbt * bt) * 0.9 * 0.75) // bd
    assert nbt * bt + nbd * bd <= free, "the two caches together must fit the free memory of the snapshot"
    # without the reservation the draft would have been promised memory the target already holds
    assert loader.kv_blocks_for(cfg, spec_d, 1, 0.75) * bd + nbt * bt > free
    # cap: never more blocks than the sequences can use; floor: one block
    monkeypatch.setattr(torch.cuda, "mem_get_info", lambda *a: (170 << 30, 180 << 30))
    assert loader.kv_blocks_for(cfg, spec_d, 1, 0.75) == 64 * 32 * 2 + 2
    assert loader.kv_blocks_for(cfg, spec_t, 1, 0.8, reserved=200 << 30) == 1
