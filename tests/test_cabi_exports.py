"""-m "not gpu": the C-ABI library builds for sm_100a, loads, and exports every symbol include/yaqs_b200.h declares
(no compute calls without a GPU)."""
import ctypes
import os
import re

from yaqs_b200.build import LIB, build
from yaqs_b200.kernels import EXPORTED_SYMBOLS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_exports_header_symbols():
    path = build(force=False)
    assert os.path.exists(path) and path == LIB
    lib = ctypes.CDLL(path)
    header = open(os.path.join(ROOT, "include", "yaqs_b200.h")).read()
    declared = set(re.findall(r"\b(yb_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found in the header"
    assert declared == set(EXPORTED_SYMBOLS)
    for sym in declared:
        assert hasattr(lib, sym), sym
    lib.yb_version.restype = ctypes.c_int
    assert lib.yb_version() >= 100


def test_sass_uses_fp64_tensor_path():
    """The contraction kernel must be on the FP64 tensor pipe (DMMA) with async-copy staging (LDGSTS)."""
    import subprocess
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    assert "DMMA" in out and "LDGSTS" in out
