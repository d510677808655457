#!/usr/bin/env python
"""Run the reference's unmodified bench/bench.py on top of ssd_b200 (the acceptance harness).

    python tools/run_reference_bench.py --ref /path/to/ssd [--synthetic] -- --size 8 --spec --k 6 --random --numseqs 8 --b 1

`--ref` is a checkout of tanishqkumar/ssd (only its bench/ directory is used; its `ssd` package is shadowed by the
compat shim).  With `--synthetic` the HF cache the harness expects ($SSD_HF_CACHE/models--meta-llama--…/snapshots/x) is
populated with config-only synthetic model dirs (random-init weights with the bigram-agreement construction of
ssd_b200/synth.py), so the harness runs without checkpoints.  `wandb` is stubbed when it is not installed."""
import argparse
import os
import runpy
import sys
import tempfile
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SYNTH = {  # HF repo directory -> (synth shape, role)
    "models--meta-llama--Llama-3.1-8B-Instruct": ("llama-3.1-8b", "target"),
    "models--meta-llama--Llama-3.3-70B-Instruct": ("llama-3.1-70b", "target"),
    "models--meta-llama--Llama-3.2-1B-Instruct": ("llama-3.2-1b", "draft"),
}


def main():
    if "--" in sys.argv:
        i = sys.argv.index("--")
        own, rest = sys.argv[1:i], sys.argv[i + 1:]
    else:
        own, rest = sys.argv[1:], []
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", required=True, help="checkout of the reference repository")
    ap.add_argument("--synthetic", action="store_true", help="populate a temporary HF cache with synthetic model dirs")
    a = ap.parse_args(own)
    bench_dir = os.path.join(a.ref, "bench")
    if not os.path.isfile(os.path.join(bench_dir, "bench.py")):
        sys.exit(f"{bench_dir}/bench.py not found")
    if a.synthetic:
        from ssd_b200 import synth
        cache = tempfile.mkdtemp(prefix="ssd_b200_hf_")
        os.environ["SSD_HF_CACHE"] = cache
        os.environ.setdefault("SSD_DATASET_DIR", os.path.join(cache, "datasets"))
        for repo, (shape, role) in SYNTH.items():
            if shape not in synth.SHAPES:
                continue
            snap = os.path.join(cache, repo, "snapshots")
            os.makedirs(snap, exist_ok=True)
            os.rename(synth.make_model_dir(snap, shape, role), os.path.join(snap, "synthetic"))
    import ssd_b200.compat as compat
    compat.install()
    try:
        import wandb  # noqa: F401
    except ImportError:
        w = types.ModuleType("wandb")
        w.init = w.log = w.finish = lambda *x, **k: None
        sys.modules["wandb"] = w
    sys.path.insert(0, bench_dir)
    sys.argv = ["bench.py"] + rest
    runpy.run_path(os.path.join(bench_dir, "bench.py"), run_name="__main__")


if __name__ == "__main__":
    main()
