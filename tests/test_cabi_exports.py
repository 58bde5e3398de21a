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


def test_workspace_bytes_is_host_arithmetic_and_matches_the_layout():
    """yb_workspace_bytes (SURVEY 8b ownership row) runs without a device and agrees with the engine's layout arithmetic; at C3 sizes
    (L = 100, chi = 128, D = 3) a trajectory needs ~205 MB, i.e. ~750 fit in 180 GB minus headroom."""
    from yaqs_b200.engine import BatchLayout
    from yaqs_b200.kernels import workspace_bytes
    L = 100
    lay = BatchLayout(L, 1, 128, [1] + [3] * (L - 1) + [1], [1] * (L + 1))
    one = workspace_bytes(L, 1, lay.cap, lay.D, 2, False)
    assert one["state"] == 2 * (16 * lay.state_elems + 4 * (L + 1))
    assert one["environments"] == 16 * lay.env_elems and one["theta"] == 16 * lay.theta_slot and one["matvec"] == 32 * lay.t_slot
    assert one["total"] == sum(v for k, v in one.items() if k != "total")
    assert 1.9e8 < one["total"] < 2.2e8
    many = workspace_bytes(L, 64, lay.cap, lay.D, 2, False)
    assert abs(many["total"] - 64 * one["total"]) < 1e-6 * many["total"]
    assert workspace_bytes(L, 1, lay.cap, lay.D, 1, True)["environments"] == 2 * one["environments"]
    small = BatchLayout(10, 1, 8, [1] + [3] * 9 + [1], [1] * 11)
    assert workspace_bytes(10, 1, small.cap, small.D, 2, False)["heev"] == 0          # 2 chi < 33: Jacobi only, no eigen-solver work
