"""End-to-end through the public API (LLM.generate, the call bench/bench.py makes): the speculative-decoding invariant
— sync-SD output == autoregressive output of the target, token for token at temperature 0 — plus the
METRICS contract, EOS / max_new_tokens truncation and prefix-cache reuse, on synthetic tiny Llama pairs."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dirs(tmp_path_factory):
    from ssd_b200 import synth
    root = str(tmp_path_factory.mktemp("models"))
    t = synth.make_model_dir(root, "llama-tiny-target", "target", seed=1, alpha=0.7, max_position_embeddings=2048)
    d = synth.make_model_dir(root, "llama-tiny-draft", "draft", seed=1, alpha=0.7, max_position_embeddings=2048)
    return t, d


def _prompts():
    g = torch.Generator().manual_seed(0)
    return [torch.randint(2, 1000, (n,), generator=g).tolist() for n in (5, 70, 33)]


def test_sd_equals_ar_and_metrics(dirs):
    from ssd_b200 import LLM, SamplingParams
    from ssd_b200.engine.llm_engine import METRICS
    t, d = dirs
    sp = SamplingParams(temperature=0.0, max_new_tokens=48, ignore_eos=True)
    K = 4
    llm = LLM(t, speculate=True, draft=d, speculate_k=K, max_num_seqs=2, max_model_len=1024, kvcache_block_size=64,
              jit_speculate=True)
    out_sd, m = llm.generate(_prompts(), sp, use_tqdm=False)
    lens = list(m["accepted_suffix_lens_with_recovery"])
    assert all(1 <= x <= K + 1 for x in lens) and len(lens) > 0
    assert m["decode_total_tokens"] == sum(lens)            # counted before truncation (verifier.py:127, step.py:163)
    assert m["decode_total_time"] > 0 and m["prefill_total_tokens"] == sum(len(p) for p in _prompts())
    assert sum(lens) / len(lens) > 1.5                      # the synthetic draft agrees with the target most of the time
    llm.exit()
    ar = LLM(t, speculate=False, max_num_seqs=2, max_model_len=1024, kvcache_block_size=64)
    out_ar, _ = ar.generate(_prompts(), sp, use_tqdm=False)
    ar.exit()
    for a, b in zip(out_sd, out_ar):
        assert len(a["token_ids"]) == 48
        assert a["token_ids"] == b["token_ids"], "speculative decoding changed the target's greedy output"


def test_eos_and_max_tokens_and_prefix_cache(dirs):
    from ssd_b200 import LLM, SamplingParams
    t, d = dirs
    llm = LLM(t, speculate=True, draft=d, speculate_k=4, max_num_seqs=1, max_model_len=1024, kvcache_block_size=64,
              jit_speculate=True)
    prompt = _prompts()[1]  # 70 tokens: one full 64-token block gets hashed
    base, _ = llm.generate([prompt], SamplingParams(temperature=0.0, max_new_tokens=30, ignore_eos=True), use_tqdm=False)
    toks = base[0]["token_ids"]
    # same prompt again: the first block is served from the prefix cache and the output must not change
    again, _ = llm.generate([prompt], SamplingParams(temperature=0.0, max_new_tokens=30, ignore_eos=True), use_tqdm=False)
    assert again[0]["token_ids"] == toks
    # declare the 10th generated token to be EOS: generation must stop right after it
    llm.scheduler.eos = toks[9]
    cut, _ = llm.generate([prompt], SamplingParams(temperature=0.0, max_new_tokens=30, ignore_eos=False), use_tqdm=False)
    first = toks.index(toks[9])
    assert cut[0]["token_ids"] == toks[:first + 1]
    # odd max_new_tokens: truncated inside an accepted suffix
    short, _ = llm.generate([prompt], SamplingParams(temperature=0.0, max_new_tokens=7, ignore_eos=True), use_tqdm=False)
    assert short[0]["token_ids"] == toks[:7]
    llm.exit()


def test_temperature_generate_runs(dirs):
    from ssd_b200 import LLM, SamplingParams
    t, d = dirs
    llm = LLM(t, speculate=True, draft=d, speculate_k=4, max_num_seqs=2, max_model_len=1024, kvcache_block_size=64,
              jit_speculate=True, seed=5)
    out, m = llm.generate(_prompts()[:2], SamplingParams(temperature=0.8, max_new_tokens=24, ignore_eos=True), use_tqdm=False)
    assert all(len(o["token_ids"]) == 24 and all(0 <= x < 1024 for x in o["token_ids"]) for o in out)
    llm.exit()
