"""CPU-only: the C-ABI library builds, loads, and exports every symbol include/spartan2_b200.h declares."""
import ctypes as C
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from spartan2_b200 import build
    path = build.build()
    L = C.CDLL(path)
    hdr = open(os.path.join(ROOT, "include", "spartan2_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(sp2_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) > 20
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_no_cpu_fallback_without_gpu():
    import torch
    import spartan2_b200 as sp
    if torch.cuda.is_available():
        return
    try:
        sp.Context(0)
    except sp.SpartanError as e:
        assert "no CPU fallback" in str(e)
    else:
        raise AssertionError("Context() must fail loudly without a CUDA device")
