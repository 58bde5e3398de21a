"""Uniform tables for the jump tests.

``numpy`` mode reproduces the reference's seeded streams bit for bit (core/random_utils.py:20-87): per trajectory one
``SeedSequence([seed, traj, "TRAJ"])`` stream for initialize/step_through and one fresh
``SeedSequence([seed, traj, timestep, "SAMP"])`` stream per measurement copy.  A jump test consumes one double
(``rng.random()``) and, if it jumps, one more (``rng.choice`` draws exactly one uniform and selects with
``searchsorted(cdf, u, side="right")``), so a table of consecutive doubles per stream is an exact replay.

``philox`` mode draws the same tables on the device with torch's Philox4x32-10 CUDA generator keyed by the run seed;
row b belongs to global trajectory index ``traj_indices[b]`` so results do not depend on how trajectories are sharded.
"""
from __future__ import annotations

import numpy as np
import torch

STREAM_TRAJECTORY = 0x5452414A  # "TRAJ"
STREAM_SAMPLE = 0x53414D50      # "SAMP"
STREAM_DISORDER = 0x4449534F    # "DISO"


def make_disorder_rng(seed: int | None) -> np.random.Generator:
    """core/random_utils.py:73-87."""
    if seed is None:
        return np.random.default_rng()
    return np.random.default_rng(np.random.SeedSequence([seed, STREAM_DISORDER]))


def numpy_tables(seed: int | None, traj_indices, n_times: int, order: int, sample_offset: int = 0):
    """(traj [B, 2*n_times], sample [B, n_times, 2] | None) float64 host arrays."""
    nb = len(traj_indices)
    traj = np.empty((nb, 2 * n_times), dtype=np.float64)
    samp = np.empty((nb, n_times, 2), dtype=np.float64) if order == 2 else None
    for row, t in enumerate(traj_indices):
        if seed is None:
            g = np.random.default_rng()
        else:
            g = np.random.default_rng(np.random.SeedSequence([seed, int(t), STREAM_TRAJECTORY]))
        traj[row] = g.random(2 * n_times)
        if samp is not None:
            for j in range(n_times):
                if seed is None:
                    gs = np.random.default_rng()
                else:
                    gs = np.random.default_rng(np.random.SeedSequence([seed, int(t), j + sample_offset, STREAM_SAMPLE]))
                samp[row, j] = gs.random(2)
    return traj, samp


def philox_tables(seed: int | None, traj_indices, n_times: int, order: int, device):
    """Device-side Philox tables; one generator per trajectory index keeps rows independent of the sharding."""
    nb = len(traj_indices)
    traj = torch.empty((nb, 2 * n_times), dtype=torch.float64, device=device)
    samp = torch.empty((nb, n_times, 2), dtype=torch.float64, device=device) if order == 2 else None
    gen = torch.Generator(device=device)
    base = int(seed) if seed is not None else int(np.random.SeedSequence().entropy % (2 ** 62))
    for row, t in enumerate(traj_indices):
        gen.manual_seed((base * 1000003 + int(t)) % (2 ** 63 - 1))
        traj[row] = torch.rand(2 * n_times, dtype=torch.float64, device=device, generator=gen)
        if samp is not None:
            samp[row] = torch.rand((n_times, 2), dtype=torch.float64, device=device, generator=gen)
    return traj, samp
