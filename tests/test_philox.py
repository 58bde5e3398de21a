"""Philox4x32-10 jump-test uniforms (yaqs_b200/csrc/yb_philox.cuh): known-answer vectors, numpy restatement, device kernel.

The reference draws from per-trajectory numpy streams (core/random_utils.py:20-87); rng_mode="philox" replaces them by a
counter-based generator addressed by (seed, global trajectory index, stream tag, jump-test index)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle.np_kernels import philox4x32_10, philox_pairs

HERE = os.path.dirname(os.path.abspath(__file__))

# Random123 kat_vectors, philox4x32 with 10 rounds: (counter, key, expected)
KAT = [
    ([0, 0, 0, 0], [0, 0], [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]),
    ([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2, [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]),
    ([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0], [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]),
]


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("philox") / "libphilox_emu.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out, os.path.join(HERE, "host_emu", "philox_emu.cpp")], check=True)
    lib = ctypes.CDLL(out)
    lib.philox_table_host.argtypes = [ctypes.c_uint64, ctypes.c_int64, ctypes.c_uint32, ctypes.c_int, ctypes.c_uint32, ctypes.c_void_p]
    return lib


@pytest.mark.parametrize("ctr,key,want", KAT)
def test_known_answer_vectors(emu, ctr, key, want):
    c = (ctypes.c_uint32 * 4)(*ctr)
    emu.philox_block_host(c, ctypes.c_uint32(key[0]), ctypes.c_uint32(key[1]))
    assert [int(x) for x in c] == want
    got = philox4x32_10(np.array([ctr], dtype=np.uint32), np.array([key], dtype=np.uint32))[0]
    assert [int(x) for x in got] == want


def test_kernel_source_matches_numpy_restatement(emu):
    for seed, traj, tag, n, b0 in [(42, 0, 0x5452414A, 17, 0), (2 ** 63 + 12345, 2 ** 40 + 7, 0x53414D50, 64, 5), (0, 4095, 1, 3, 2 ** 32 - 2)]:
        out = np.empty(2 * n)
        emu.philox_table_host(seed, traj, tag, n, b0, out.ctypes.data_as(ctypes.c_void_p))
        blocks = (np.arange(b0, b0 + n, dtype=np.uint64)) & np.uint64(0xFFFFFFFF)
        want = philox_pairs(seed, traj, tag, blocks).reshape(-1)
        assert np.array_equal(out, want)
        assert out.min() >= 0.0 and out.max() < 1.0


def test_uniformity_and_stream_independence():
    u = philox_pairs(7, 3, 0x5452414A, np.arange(200000, dtype=np.uint64)).reshape(-1)
    assert abs(u.mean() - 0.5) < 4 * np.sqrt(1 / 12 / u.size)
    hist, _ = np.histogram(u, bins=64, range=(0, 1))
    chi2 = float(((hist - u.size / 64) ** 2 / (u.size / 64)).sum())
    assert chi2 < 63 + 6 * np.sqrt(2 * 63)
    v = philox_pairs(7, 4, 0x5452414A, np.arange(200000, dtype=np.uint64)).reshape(-1)
    w = philox_pairs(7, 3, 0x53414D50, np.arange(200000, dtype=np.uint64)).reshape(-1)
    for other in (v, w):
        assert abs(np.corrcoef(u, other)[0, 1]) < 5 / np.sqrt(u.size)


@pytest.mark.gpu
def test_device_tables_bit_exact_and_shard_invariant():
    import torch

    from yaqs_b200.kernels import CudaKernels
    from yaqs_b200.rng import STREAM_SAMPLE, STREAM_TRAJECTORY, philox_tables
    kern = CudaKernels("cuda:0")
    seed, n_times = 987654321987, 37
    idx = [0, 1, 5, 4095, 2 ** 33 + 3]
    tu, su = philox_tables(kern, seed, idx, n_times, 2, "cuda:0", sample_offset=2)
    torch.cuda.synchronize()
    tu, su = tu.cpu().numpy(), su.cpu().numpy()
    for row, t in enumerate(idx):
        assert np.array_equal(tu[row], philox_pairs(seed, t, STREAM_TRAJECTORY, np.arange(n_times, dtype=np.uint64)).reshape(-1))
        assert np.array_equal(su[row], philox_pairs(seed, t, STREAM_SAMPLE, np.arange(2, 2 + n_times, dtype=np.uint64)))
    tu2, _ = philox_tables(kern, seed, idx[2:4], n_times, 1, "cuda:0")          # another batch composition: same rows
    assert np.array_equal(tu2.cpu().numpy(), tu[2:4])
