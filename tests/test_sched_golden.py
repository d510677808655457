"""Scheduler / BlockManager / Sequence against the reference's own classes: the scripted workloads of
oracle/sched_driver.py were run through the unmodified ssd/engine/{scheduler,block_manager,sequence}.py
(oracle/gen_sched_golden.py); here the same driver runs ssd_b200.engine and every recorded state must be identical —
block tables and free-list order included, because the block tables are hot-path inputs of the kernels."""
import gzip
import json
import os

import pytest

from oracle import sched_driver

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name", sorted(sched_driver.SCENARIOS))
def test_bookkeeping_matches_reference(name):
    from ssd_b200.engine.scheduler import Scheduler
    from ssd_b200.engine.sequence import Sequence
    from ssd_b200.sampling_params import SamplingParams

    with gzip.open(os.path.join(GOLD, f"sched_{name}.json.gz")) as f:
        want = json.loads(f.read())
    old_bs = Sequence.block_size
    try:
        got = sched_driver.run(name, Scheduler, Sequence, SamplingParams)
    finally:
        Sequence.block_size = old_bs
    assert got["eos"] == want["eos"]
    assert len(got["steps"]) == len(want["steps"])
    for i, (g, w) in enumerate(zip(got["steps"], want["steps"])):
        assert g["is_prefill"] == w["is_prefill"] and g["batch"] == w["batch"], f"step {i}: schedule differs"
        assert g.get("suffix_lens") == w.get("suffix_lens")
        for j, (gs, ws) in enumerate(zip(g["after"]["seqs"], w["after"]["seqs"])):
            assert gs == ws, f"step {i}, sequence {j}: {gs} != {ws}"
        for key in ("target", "draft", "waiting", "running"):
            gk, wk = g["after"][key], w["after"][key]
            if isinstance(gk, dict):  # the number of registered hashes may differ: see test_block_sealed_is_the_block_that_completed
                gk, wk = {k: v for k, v in gk.items() if k != "hashed"}, {k: v for k, v in wk.items() if k != "hashed"}
            assert gk == wk, f"step {i}: {key} differs"
    assert got["outputs"] == want["outputs"]


def test_block_sealed_is_the_block_that_completed():
    """Deliberate divergence from a reference quirk.  When an accepted suffix fills block i and spills into block i+1,
    Scheduler._finalize_block (ssd/engine/scheduler.py:243-250) hashes the tokens of block i but registers the hash on
    `block_table[-1]` with the prefix of `block_table[-2]` — the page of block i+1 — so a later prompt with the same
    prefix would be served a page holding the wrong tokens' KV (and block i is re-finalised every step because its own
    hash stays -1).  ssd_b200 seals the block that actually completed; everything else in the bookkeeping is identical
    (test above)."""
    from types import SimpleNamespace

    from ssd_b200.engine.block_manager import BlockManager
    from ssd_b200.engine.scheduler import Scheduler
    from ssd_b200.engine.sequence import Sequence
    from ssd_b200.sampling_params import SamplingParams

    old_bs = Sequence.block_size
    Sequence.block_size = 4
    try:
        cfg = SimpleNamespace(max_num_seqs=2, max_num_batched_tokens=4096, max_model_len=256, eos=-1, speculate=True,
                              speculate_k=3, kvcache_block_size=4, num_kvcache_blocks=8)
        sched = Scheduler(cfg, SimpleNamespace(num_kvcache_blocks=8))
        a = Sequence([10, 11, 12], SamplingParams(temperature=0.0, max_new_tokens=32, ignore_eos=True))
        sched.add(a)
        batch, is_prefill = sched.schedule()
        assert is_prefill and batch == [a]
        a.recovery_token_id, a.num_cached_tokens, a.num_draft_cached_tokens = 13, 3, 3
        batch, is_prefill = sched.schedule()
        assert not is_prefill
        sched.postprocess_speculate([a], [[13, 14, 15]], [16])  # tokens 10..15: block 0 full, block 1 half full
        for m, table in ((sched.block_manager, a.block_table), (sched.draft_block_manager, a.draft_block_table)):
            h0 = BlockManager.compute_hash([10, 11, 12, 13])
            assert m.blocks[table[0]].digest == h0 and m.hash_to_block_id[h0] == table[0]
            assert m.blocks[table[1]].digest == -1
        b = Sequence([10, 11, 12, 13, 99], SamplingParams(temperature=0.0, max_new_tokens=4, ignore_eos=True))
        sched.add(b)
        batch, is_prefill = sched.schedule()
        assert is_prefill and batch == [b]
        assert b.block_table[0] == a.block_table[0] and b.num_cached_tokens == 4  # the page that really holds 10..13
        assert b.draft_block_table[0] == a.draft_block_table[0] and b.num_draft_cached_tokens == 4
    finally:
        Sequence.block_size = old_bs
