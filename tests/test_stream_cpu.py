"""stream_callback contract of LLM.generate (reference: engine/llm_engine.py:356-369, used by bench/chat.py:96-102):
called as callback(seq_id, new_token_ids) after every engine step for running sequences and once more with the tail of
a finished sequence; the concatenation per sequence equals the returned token_ids.  CPU only (device runner doubled)."""
import types

import pytest

from tests.test_reference_harness_cpu import FakeRunner, _next


@pytest.mark.parametrize("speculate", [True, False])
def test_stream_callback_concatenates_to_the_outputs(tmp_path, monkeypatch, speculate):
    from ssd_b200 import LLM, SamplingParams, synth
    import ssd_b200.loader as loader

    target = synth.make_model_dir(str(tmp_path), "llama-3.1-8b", "target")
    draft = synth.make_model_dir(str(tmp_path), "llama-3.2-1b", "draft")

    def fake_build_runner(config, tp_size=1, tp_rank=0, device=None, finalize=True):
        config.num_kvcache_blocks = 64
        return FakeRunner(config.speculate_k), types.SimpleNamespace(num_kvcache_blocks=64)

    monkeypatch.setattr(loader, "build_runner", fake_build_runner)
    llm = LLM(target, speculate=speculate, draft=draft, speculate_k=4, num_gpus=1, max_num_seqs=2, max_model_len=1024)
    prompts = [[5, 6, 7], [100, 200], [42]]
    sps = [SamplingParams(temperature=0.0, max_new_tokens=n, ignore_eos=True) for n in (17, 9, 30)]
    streamed, calls = {}, []

    def on_tokens(seq_id, new_ids):
        assert len(new_ids) > 0
        streamed.setdefault(seq_id, []).extend(new_ids)
        calls.append(seq_id)

    outs, _ = llm.generate(prompts, sps, use_tqdm=False, stream_callback=on_tokens)
    ids = sorted(streamed)
    assert len(ids) == 3
    for sid, p, o, sp in zip(ids, prompts, outs, sps):
        want, cur = [], p[-1]
        for _ in range(sp.max_new_tokens):
            cur = _next(cur)
            want.append(cur)
        assert o["token_ids"] == want
        assert streamed[sid] == want
    assert len(calls) > 3  # incremental, not one call per sequence
