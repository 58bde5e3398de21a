"""The keep decision of every SVD on the path -- truncate() (core/linalg/svd_utils.py:22-104) -- pinned on the CPU: the oracle's
restatement AND the device function's own source (yaqs_b200/csrc/yb_truncate.cuh, built for the host with g++) against the known
answers of the reference's tests (tests/core/linalg/test_svd_utils.py:20-85), and against each other on random spectra in all four
modes.  Keep counts are integers: the bar is equality."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle.tjm_oracle import truncate

HERE = os.path.dirname(os.path.abspath(__file__))
MODES = {"discarded_weight": 0, "relative": 1, "hard_cutoff": 2, "relative_discarded_weight": 3}

# (spectrum, mode, threshold, max_bond_dim, min_keep, expected keep) -- the cases of the reference's own tests, file:line in the comment
KNOWN = [
    ([10.0, 3.0, 1.0, 0.5], "discarded_weight", 10.0, None, 1, 2),                     # test_svd_utils.py:20-25
    ([2.0, 1.0, 0.4], "relative", 0.45, None, 1, 2),                                   # :28-32
    ([0.0, 1.0], "relative", 0.1, None, 1, 1),                                         # :33-35 (s_max == 0 -> min_keep)
    ([5.0, 2.0, 0.5, 0.1], "hard_cutoff", 0.2, None, 1, 3),                            # :38-42
    ([5.0, 2.0, 0.5, 0.1], "hard_cutoff", 0.2, 2, 1, 2),                               # :43-44
    ([10.0, 1.0, 0.1, 0.01], "relative_discarded_weight", 1e-3, None, 1, 2),           # :54-60
    ([10.0, 1.0, 0.1, 0.01], "relative_discarded_weight", 0.02, None, 1, 1),           # :61-62
    ([3.0, 2.0, 0.0, 0.0], "relative_discarded_weight", 0.0, None, 1, 2),              # :74-77
    ([0.0, 0.0, 0.0, 0.0], "relative_discarded_weight", 0.1, None, 1, 1),              # :78-79
    ([5.0, 4.0, 3.0, 2.0], "relative_discarded_weight", 1.0, None, 2, 2),              # :80-81
    ([5.0, 4.0, 3.0, 2.0], "relative_discarded_weight", 1.0, 1, 1, 1),                 # :82
]


@pytest.fixture(scope="module")
def device_truncate(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("emu") / "truncate_emu.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out, os.path.join(HERE, "host_emu", "truncate_emu.cpp")], check=True)
    lib = ctypes.CDLL(out)
    lib.truncate_keep_host.restype = ctypes.c_int
    lib.truncate_keep_host.argtypes = [ctypes.POINTER(ctypes.c_double), ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_int]

    def call(s, mode, threshold, max_bond, min_keep):
        a = np.ascontiguousarray(s, dtype=np.float64)
        return int(lib.truncate_keep_host(a.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), int(a.size), MODES[mode], float(threshold),
                                          0 if max_bond is None else int(max_bond), int(min_keep)))
    return call


@pytest.mark.parametrize("case", KNOWN)
def test_known_answers_of_the_reference(device_truncate, case):
    s, mode, thr, mb, mk, expected = case
    assert truncate(np.array(s), mode, thr, mb, mk) == expected
    assert device_truncate(s, mode, thr, mb, mk) == expected


def test_relative_discarded_weight_is_scale_invariant(device_truncate):
    """test_svd_utils.py:65-71."""
    s = np.array([4.0, 2.0, 0.5, 0.1])
    ref = truncate(s, "relative_discarded_weight", 0.05, None, 1)
    for scale in (3.7, 1e200, 1e-200):
        assert truncate(scale * s, "relative_discarded_weight", 0.05, None, 1) == ref
        assert device_truncate(scale * s, "relative_discarded_weight", 0.05, None, 1) == ref


def test_unknown_mode_raises_in_the_oracle():
    """test_svd_utils.py:47-51 (the device function only ever sees the four mode codes)."""
    with pytest.raises(ValueError, match="Unknown truncation mode"):
        truncate(np.ones(3), "invalid", 1.0, None)


@pytest.mark.parametrize("mode", sorted(MODES))
def test_device_source_equals_oracle_on_random_spectra(device_truncate, mode):
    rng = np.random.default_rng(MODES[mode])
    for trial in range(400):
        n = int(rng.integers(1, 40))
        s = np.sort(rng.random(n) * 10.0 ** rng.uniform(-8, 1, n))[::-1].copy()
        if trial % 7 == 0:
            s[n // 2:] = 0.0                                   # rank-deficient tails
        thr = float(10.0 ** rng.uniform(-14, 0.5))
        mb = None if trial % 3 else int(rng.integers(1, n + 3))
        mk = int(rng.integers(1, 3))
        assert device_truncate(s, mode, thr, mb, mk) == truncate(s, mode, thr, mb, mk), (mode, trial, s, thr, mb, mk)
