"""Drop-in shim: after `ssd_b200.compat.install()`, the imports bench/bench.py performs
(`from ssd import LLM, SamplingParams`, `from ssd.engine.llm_engine import METRICS`, `import ssd.paths`,
bench_helpers' `from ssd.paths import DATASET_PATHS, HF_CACHE_DIR, EAGLE3_*`) resolve to this package."""
import sys
import types


def install() -> None:
    import ssd_b200
    from ssd_b200 import config, llm, paths, sampling_params
    from ssd_b200.engine import llm_engine, scheduler, sequence

    root = types.ModuleType("ssd")
    root.LLM, root.SamplingParams, root.Config = llm.LLM, sampling_params.SamplingParams, config.Config
    root.Sequence, root.SequenceStatus = sequence.Sequence, sequence.SequenceStatus
    root.LLMEngine = llm_engine.LLMEngine
    root.__path__ = []
    eng = types.ModuleType("ssd.engine")
    eng.__path__ = []
    mods = {
        "ssd": root, "ssd.paths": paths, "ssd.config": config, "ssd.sampling_params": sampling_params, "ssd.llm": llm,
        "ssd.engine": eng, "ssd.engine.llm_engine": llm_engine, "ssd.engine.scheduler": scheduler,
        "ssd.engine.sequence": sequence,
    }
    root.paths, root.config, root.engine = paths, config, eng
    eng.llm_engine, eng.scheduler, eng.sequence = llm_engine, scheduler, sequence
    for k, v in mods.items():
        sys.modules[k] = v
