"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol that
include/ssdk.h declares; calls that need a GPU fail loudly (no silent CPU fallback)."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _declared_symbols():
    text = (ROOT / "include" / "ssdk.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ssdk_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_bound_and_exported():
    from ssd_b200 import lib
    declared = _declared_symbols()
    assert len(declared) >= 25
    assert sorted(lib.SIGNATURES) == declared, "ssd_b200/lib.py and include/ssdk.h disagree"
    so = lib.load()
    for name in declared:
        assert hasattr(so, name), f"libssdk.so does not export {name}"
    assert so.ssdk_abi_version() == 1


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from ssd_b200 import lib, ops
    from ssd_b200.runner import ModelSpec, PairRunner
    with pytest.raises(RuntimeError):
        PairRunner(ModelSpec(128, 1, 2, 1, 64, 256, 512), None, spec_k=0)
    with pytest.raises(RuntimeError):
        ops.sample(torch.zeros(1, 8, dtype=torch.bfloat16), torch.zeros(1))
    so = lib.load()
    h = ctypes.c_void_p()
    cfg = lib.ModelCfg(128, 1, 2, 1, 64, 256, 512, 0, 1e-5, 256, 1, 0)
    rt = lib.RuntimeCfg(0, 1, 256, 1, 0, 0, 0, 0)
    assert so.ssdk_create(ctypes.byref(cfg), None, ctypes.byref(rt), ctypes.byref(h)) != 0
    assert b"CUDA" in so.ssdk_last_error() or b"device" in so.ssdk_last_error()


def test_product_never_imports_oracle():
    for p in (ROOT / "ssd_b200").rglob("*.py"):
        src = p.read_text()
        assert "import oracle" not in src and "from oracle" not in src, f"{p} imports the oracle"
