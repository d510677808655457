"""Bookkeeping-level driver for differential tests of Scheduler / BlockManager / Sequence.

TEST INFRASTRUCTURE ONLY.  The same scripted workload is pushed through the reference's own classes
(oracle/gen_sched_golden.py, in the build container) and through ssd_b200.engine (tests/test_sched_golden.py); the model
is replaced by a deterministic token chain with seeded accept lengths, exactly where LLMEngine.step
(engine/llm_engine.py:186-236) would call ModelRunner / Speculator / Verifier:

  prefill, speculative   SpecDecodeStep.prefill (engine/step.py:73-89): recovery token set, cached counters = prompt length
  prefill, autoregressive AutoRegressiveStep.step -> Scheduler.postprocess(..., is_prefill=True) (step.py:36-47)
  decode, speculative    Scheduler.postprocess_speculate(seqs, [recovery + accepted], next_recovery) (step.py:148-153)
  decode, autoregressive Scheduler.postprocess(seqs, tokens, False)

After every engine step the complete observable state is recorded: per sequence (status, counters, both block tables,
recovery token), per block manager (free list order, used set, number of hashed blocks), queue orders.
"""
from __future__ import annotations

import random
from types import SimpleNamespace

V = 50021


def nxt(tok: int) -> int:
    return (tok * 31 + 7) % V


SCENARIOS = {
    # name: engine config + requests (prompt spec, max_new_tokens, ignore_eos)
    "spec_preempt_prefix": dict(speculate=True, K=4, block_size=16, num_blocks=14, draft_blocks=14, max_num_seqs=3,
                                max_model_len=4096, seed=1,
                                requests=[(("shared", 40, 3), 30, True), (("shared", 40, 9), 45, True), (("rand", 21), 60, False),
                                          (("shared", 40, 5), 17, True), (("rand", 33), 25, False), (("rand", 5), 64, True)]),
    # K=6: EOS inside a suffix, max_new_tokens cutting a suffix, and one always-accepting request that runs exactly into
    # max_model_len (47 + 5 * 7 = 82: the last step still fits the look-ahead, its suffix reaches the limit)
    "spec_truncation": dict(speculate=True, K=6, block_size=16, num_blocks=40, draft_blocks=40, max_num_seqs=2,
                            max_model_len=82, seed=2,
                            requests=[(("rand", 47), 200, True, "full"), (("rand", 30), 13, True), (("shared", 32, 2), 10, False),
                                      (("shared", 32, 4), 29, False), (("rand", 9), 3, True), (("rand", 11), 1, True)]),
    "spec_k1_big_blocks": dict(speculate=True, K=1, block_size=256, num_blocks=6, draft_blocks=5, max_num_seqs=4,
                               max_model_len=2048, seed=3,
                               requests=[(("rand", 250), 20, True), (("rand", 255), 9, True), (("rand", 300), 300, True),
                                         (("shared", 256, 3), 12, True), (("shared", 256, 8), 30, True)]),
    "ar_preempt_prefix": dict(speculate=False, K=0, block_size=16, num_blocks=12, draft_blocks=0, max_num_seqs=3,
                              max_model_len=4096, seed=4,
                              requests=[(("shared", 32, 3), 40, True), (("shared", 32, 9), 30, False), (("rand", 15), 50, True),
                                        (("rand", 47), 33, False), (("shared", 32, 1), 16, True)]),
}


def _prompt(rng, spec):
    if spec[0] == "rand":
        return [rng.randrange(2, V) for _ in range(spec[1])]
    base = [(i * 7919 + 13) % V for i in range(spec[1])]  # shared prefix (prefix-cache hits)
    return base + [rng.randrange(2, V) for _ in range(spec[2])]


def _snapshot(sched, seqs):
    def bm(m):
        if m is None:
            return None
        return {"free": list(m.free_block_ids), "used": sorted(m.used_block_ids), "hashed": len(m.hash_to_block_id),
                "ref": [getattr(b, "ref_count", getattr(b, "refs", None)) for b in m.blocks]}
    idx = {id(s): i for i, s in enumerate(seqs)}
    return {
        "seqs": [{"status": s.status.name, "num_tokens": s.num_tokens, "num_prompt_tokens": s.num_prompt_tokens,
                  "num_cached_tokens": s.num_cached_tokens, "num_draft_cached_tokens": s.num_draft_cached_tokens,
                  "block_table": list(s.block_table), "draft_block_table": list(s.draft_block_table),
                  "recovery_token_id": s.recovery_token_id, "last_spec_step_accepted_len": s.last_spec_step_accepted_len,
                  "last_token": s.last_token} for s in seqs],
        "target": bm(sched.block_manager), "draft": bm(getattr(sched, "draft_block_manager", None)),
        "waiting": [idx[id(s)] for s in sched.waiting], "running": [idx[id(s)] for s in sched.running],
    }


def run(name: str, Scheduler, Sequence, SamplingParams, max_steps: int = 2000) -> dict:
    sc = SCENARIOS[name]
    rng = random.Random(sc["seed"])
    Sequence.block_size = sc["block_size"]
    prompts = [_prompt(rng, r[0]) for r in sc["requests"]]
    # an EOS that the third request's chain reaches after 11 tokens (others may or may not meet it)
    eos = prompts[2][-1]
    for _ in range(11):
        eos = nxt(eos)
    K = sc["K"]
    cfg = SimpleNamespace(max_num_seqs=sc["max_num_seqs"], fan_out_list=None, fan_out_list_miss=None, draft_async=False,
                          max_num_batched_tokens=16384, max_model_len=sc["max_model_len"], eos=eos, speculate=sc["speculate"],
                          async_fan_out=3, speculate_k=K, kvcache_block_size=sc["block_size"], verbose=False,
                          num_kvcache_blocks=sc["num_blocks"], model="unused")
    dcfg = SimpleNamespace(num_kvcache_blocks=sc["draft_blocks"], kvcache_block_size=sc["block_size"])
    sched = Scheduler(cfg, dcfg if sc["speculate"] else None)
    seqs = []
    policy = {}
    for p, r in zip(prompts, sc["requests"]):
        s = Sequence(p, SamplingParams(temperature=0.0, max_new_tokens=r[1], ignore_eos=r[2]))
        policy[id(s)] = r[3] if len(r) > 3 else "random"
        seqs.append(s)
        sched.add(s)
    idx = {id(s): i for i, s in enumerate(seqs)}
    steps = []
    while not sched.is_finished():
        assert len(steps) < max_steps, "scheduler did not terminate"
        batch, is_prefill = sched.schedule()
        rec = {"is_prefill": bool(is_prefill), "batch": [idx[id(s)] for s in batch]}
        if not batch:
            raise AssertionError("scheduler returned an empty batch")
        if is_prefill:
            if sc["speculate"]:
                for s in batch:
                    s.recovery_token_id = nxt(s.last_token)
                    s.num_cached_tokens = s.num_prompt_tokens
                    s.num_draft_cached_tokens = s.num_prompt_tokens
            else:
                sched.postprocess(batch, [nxt(s.last_token) for s in batch], True)
        elif sc["speculate"]:
            suffixes, recs = [], []
            for s in batch:
                n = rng.randint(0, K)
                if policy[id(s)] == "full":
                    n = K
                suf = [s.recovery_token_id]
                for _ in range(n):
                    suf.append(nxt(suf[-1]))
                suffixes.append(suf)
                recs.append(nxt(suf[-1]))
            rec["suffix_lens"] = [len(x) for x in suffixes]
            sched.postprocess_speculate(batch, suffixes, recs)
        else:
            sched.postprocess(batch, [nxt(s.last_token) for s in batch], False)
        rec["after"] = _snapshot(sched, seqs)
        steps.append(rec)
    return {"scenario": name, "eos": eos, "steps": steps,
            "outputs": [list(s.token_ids) for s in seqs]}
