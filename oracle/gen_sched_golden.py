"""Generate tests/golden/sched_*.json.gz by running the UNMODIFIED reference Scheduler / BlockManager / Sequence
(ssd/engine/{scheduler,block_manager,sequence}.py) through oracle/sched_driver.py.  Build container only.

    python oracle/gen_sched_golden.py

The only stub besides the import shims of gen_golden.py is the tokenizer lookup in Scheduler.__init__
(scheduler.py:32), which needs a checkpoint directory."""
import gzip
import json
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))

from oracle import gen_golden, sched_driver  # noqa: E402


def main():
    gen_golden.import_reference()
    import ssd.engine.scheduler as rs
    from ssd.engine.sequence import Sequence
    from ssd.sampling_params import SamplingParams

    class _Tok:
        @staticmethod
        def from_pretrained(*a, **k):
            return None

    rs.AutoTokenizer = _Tok
    out = REPO / "tests" / "golden"
    for name in sched_driver.SCENARIOS:
        trace = sched_driver.run(name, rs.Scheduler, Sequence, SamplingParams)
        n_pre = sum(1 for s in trace["steps"] if s["is_prefill"])
        path = out / f"sched_{name}.json.gz"
        with gzip.GzipFile(path, "wb", mtime=0) as f:
            f.write(json.dumps(trace, separators=(",", ":")).encode())
        n_preempt = sum(1 for a, b in zip(trace["steps"], trace["steps"][1:])
                        for x, y in zip(a["after"]["seqs"], b["after"]["seqs"]) if x["status"] == "RUNNING" and y["status"] == "WAITING")
        print(f"{name}: {len(trace['steps'])} steps ({n_pre} prefill, {n_preempt} preemptions), {path.stat().st_size} bytes")


if __name__ == "__main__":
    main()
