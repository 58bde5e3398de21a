"""Numerics of the divide & conquer tridiagonal eigen-solver (yaqs_b200/csrc/yb_stedc.cuh) checked against LAPACK on the CPU.

The kernel body is written as barrier-separated phases; tests/host_emu/stedc_emu.cpp compiles the SAME source with g++, running
the threads of every phase one after the other.  This pins the algorithm (tearing, leaf QL, deflation, secular equation,
Gu/Eisenstat vectors, merge product); the `-m gpu` tests (test_heev_*) check the parallel execution on the device."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("stedc") / "libstedc_emu.so")
    src = os.path.join(HERE, "host_emu", "stedc_emu.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out, src], check=True)
    return ctypes.CDLL(out)


def _solve(lib, d, e):
    k = len(d)
    d = np.ascontiguousarray(d, dtype=np.float64).copy()
    e = np.ascontiguousarray(e, dtype=np.float64).copy()
    z = np.zeros(k * k)
    rc = lib.stedc_host(d.ctypes.data_as(ctypes.c_void_p), e.ctypes.data_as(ctypes.c_void_p), k, z.ctypes.data_as(ctypes.c_void_p), 256)
    return rc, d, z.reshape(k, k).T


def _check(lib, d, e, tol=2e-13):
    k = len(d)
    t = np.diag(d) + np.diag(e[:k - 1], 1) + np.diag(e[:k - 1], -1)
    rc, lam, q = _solve(lib, d, e)
    assert rc == 0
    nrm = max(np.max(np.abs(t)), 1e-300)
    assert np.max(np.abs(t @ q - q * lam[None, :])) <= tol * nrm
    assert np.max(np.abs(q.T @ q - np.eye(k))) <= tol
    assert np.max(np.abs(np.sort(lam) - np.linalg.eigvalsh(t))) <= tol * nrm


def _gram_tridiagonal(rng, k, decay):
    import scipy.linalg as sl
    a = rng.standard_normal((k, k)) + 1j * rng.standard_normal((k, k))
    u, _, vh = np.linalg.svd(a)
    w = (u * np.logspace(0, -decay, k)) @ vh
    h = sl.hessenberg(w @ w.conj().T)
    return np.real(np.diag(h)).copy(), np.concatenate([np.abs(np.diag(h, -1)), [0.0]])


@pytest.mark.parametrize("k", [1, 2, 3, 31, 32, 33, 64, 65, 97, 128, 200, 256, 300, 512])
def test_random_and_structured(emu, k):
    rng = np.random.default_rng(k)
    _check(emu, rng.standard_normal(k), rng.standard_normal(k))
    _check(emu, rng.standard_normal(k), -np.abs(rng.standard_normal(k)))
    _check(emu, np.ones(k), np.zeros(k))
    _check(emu, np.zeros(k), np.zeros(k))
    _check(emu, 2 * np.ones(k), -np.ones(k))
    _check(emu, np.abs(np.arange(k) - k // 2).astype(float), np.ones(k))                  # Wilkinson: pairs of close eigenvalues
    _check(emu, 1e-150 * rng.standard_normal(k), 1e-150 * rng.standard_normal(k))
    _check(emu, 1e150 * rng.standard_normal(k), 1e150 * rng.standard_normal(k))
    dd = np.ones(k)
    dd[::2] += 1e-15
    _check(emu, dd, 1e-8 * np.ones(k))
    _check(emu, np.tile([1.0, 2.0, 3.0, 4.0, 5.0], k // 5 + 1)[:k], np.tile([1e-9, 1, 1e-9, 1, 1e-3], k // 5 + 1)[:k])


@pytest.mark.parametrize("k", [64, 128, 256])
@pytest.mark.parametrize("decay", [0, 3, 8, 16])
def test_gram_spectra(emu, k, decay):
    """Tridiagonal forms of Gram matrices with Schmidt-like graded spectra (what the SVD route feeds the solver)."""
    d, e = _gram_tridiagonal(np.random.default_rng(100 * k + decay), k, decay)
    _check(emu, d, e)


def test_non_finite_input_is_flagged(emu):
    rc, _, _ = _solve(emu, np.array([1.0, np.nan, 2.0] * 20), np.ones(60))
    assert rc == 1
