"""CPU checks of bench.py's bookkeeping: the HBM step-roofline formula reproduces BASELINE.md §3, and the reference arm
(`--impl reference`, the oracle port on host cores) runs end to end on a tiny workload and prints the required keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bytes_per_step_matches_baseline_md():
    sys.path.insert(0, ROOT)
    import bench
    peak = 6571.2e9
    rows = {("llama-3.1-8b", 1): 4.94, ("llama-3.1-70b", 1): 23.82, ("llama-3.1-70b", 2): 13.23, ("llama-3.1-70b", 4): 7.94,
            ("llama-3.1-70b", 8): 5.29}
    for (t, tp), ms in rows.items():
        got = bench.bytes_per_step(t, "llama-3.2-1b", 6, 384, tp) / peak * 1e3
        assert abs(got - ms) < 0.02, (t, tp, got, ms)
    q = bench.bytes_per_step("qwen3-32b", "qwen3-0.6b", 6, 384, 4) / peak * 1e3
    assert abs(q - 3.75) < 0.02


def test_reference_arm_prints_contract_keys():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "tiny",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["impl"] == "reference" and d["unit"] == "tokens/s" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
