"""Index arithmetic of the gate-MPO site kernel (yaqs_b200/csrc/yb_mpo_apply.cuh) checked on the CPU: the element function the CUDA
kernel k_mpo_site_apply runs per thread is compiled with g++ (tests/host_emu/mpo_apply_emu.cpp) and compared with the definition
``A'[s,(l a),(r b)] = sum_t W[s,t,a,b] A[t,l,r]`` (mpo_utils.py:27-56) and with the numpy kernel twin the host engine uses.  The
`-m gpu` test test_mpo_site_apply_matches_twin checks the launch (grid, masks, strides) of the same source on the device."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle.np_kernels import NumpyKernels
from yaqs_b200.kernels import Buf, Ext

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("mpo_apply") / "libmpo_apply_emu.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out, os.path.join(HERE, "host_emu", "mpo_apply_emu.cpp")], check=True)
    lib = ctypes.CDLL(out)
    lib.mpo_site_apply_host.argtypes = [ctypes.c_int] * 3 + [ctypes.c_longlong] * 2 + [ctypes.c_void_p] * 3
    return lib


@pytest.mark.parametrize("d,dl,dr,l,r", [(2, 1, 2, 3, 5), (2, 4, 4, 7, 2), (2, 2, 1, 1, 6), (2, 1, 1, 4, 4), (4, 3, 2, 2, 3), (2, 2, 2, 1, 1)])
def test_element_function_matches_definition_and_twin(emu, d, dl, dr, l, r):
    rng = np.random.default_rng(d * 1000 + dl * 100 + dr * 10 + l + r)
    w = rng.standard_normal((d, d, dl, dr)) + 1j * rng.standard_normal((d, d, dl, dr))
    a = rng.standard_normal((d, l, r)) + 1j * rng.standard_normal((d, l, r))
    out = np.full(d * l * dl * r * dr, np.nan + 0j)
    wc, ac = np.ascontiguousarray(w), np.ascontiguousarray(a)
    emu.mpo_site_apply_host(d, dl, dr, l, r, wc.ctypes.data_as(ctypes.c_void_p), ac.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p))
    want = np.einsum("stab,tlr->slarb", w, a).reshape(-1)
    assert np.max(np.abs(out - want)) < 1e-13
    # the reference's own formulation: tensordot over the input leg, transpose (0, 3, 1, 4, 2), fuse
    theta = np.tensordot(w, a, axes=([1], [0]))
    assert np.max(np.abs(out - theta.transpose(0, 3, 1, 4, 2).reshape(-1))) < 1e-13
    # the numpy twin of the kernel entry point (what the host engine runs)
    ext = torch.tensor([[l, r]], dtype=torch.int32)
    dst = torch.zeros(out.size, dtype=torch.complex128)
    NumpyKernels().mpo_site_apply(1, None, Ext(ext, 0, 2), Ext(ext, 1, 2), d, dl, dr, torch.from_numpy(wc.reshape(-1)), Buf(torch.from_numpy(ac.reshape(-1)), 0, a.size),
                                  Buf(dst, 0, out.size), out.size // d)
    assert np.max(np.abs(dst.numpy() - out)) < 1e-13


@pytest.fixture(scope="module")
def wapply_emu(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("wapply") / "libwapply_emu.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out, os.path.join(HERE, "host_emu", "wapply_emu.cpp")], check=True)
    lib = ctypes.CDLL(out)
    lib.wapply_onepass_host.argtypes = [ctypes.c_int] * 6 + [ctypes.c_void_p] * 3
    return lib


@pytest.mark.parametrize("dims", [(4, 4, 3, 3), (2, 2, 5, 4), (2, 2, 1, 1), (4, 4, 5, 5), (4, 4, 4, 3), (1, 1, 3, 3)])
@pytest.mark.parametrize("l,r", [(1, 1), (7, 5), (16, 3)])
def test_wapply_onepass_element_matches_definition(wapply_emu, dims, l, r):
    """A/B kernel k_wapply_onepass (yaqs_b200/csrc/yb_wapply.cuh): out[x,l,y,r] = sum_(u,v) wt[x,u,y,v] in[u,l,v,r] with every input
    element read once; the element function compiled with g++ against the einsum of the definition and the numpy kernel twin, sparse
    MPO tensors included (zero entries are skipped)."""
    nx, nu, ny, nv = dims
    rng = np.random.default_rng(nx * 1000 + ny * 100 + l * 10 + r)
    w = rng.standard_normal((nx, nu, ny, nv)) + 1j * rng.standard_normal((nx, nu, ny, nv))
    w.reshape(-1)[::3] = 0
    src = rng.standard_normal((nu, l, nv, r)) + 1j * rng.standard_normal((nu, l, nv, r))
    out = np.full(nx * l * ny * r, np.nan + 0j)
    wc, sc = np.ascontiguousarray(w), np.ascontiguousarray(src)
    wapply_emu.wapply_onepass_host(nx, nu, ny, nv, l, r, wc.ctypes.data_as(ctypes.c_void_p), sc.ctypes.data_as(ctypes.c_void_p),
                                   out.ctypes.data_as(ctypes.c_void_p))
    want = np.einsum("xuyv,ulvr->xlyr", w, src).reshape(-1)
    assert np.max(np.abs(out - want)) < 1e-13
    ext = torch.tensor([[l, r]], dtype=torch.int32)
    dst = torch.zeros(out.size, dtype=torch.complex128)
    NumpyKernels().wapply(1, None, Ext(ext, 0, 2), Ext(ext, 1, 2), nx, nu, ny, nv, torch.from_numpy(wc.reshape(-1)), Buf(torch.from_numpy(sc.reshape(-1)), 0, src.size),
                          Buf(dst, 0, out.size), l * r)
    assert np.max(np.abs(dst.numpy() - out)) < 1e-13
