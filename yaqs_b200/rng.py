"""Uniform tables for the jump tests.

``numpy`` mode reproduces the reference's seeded streams bit for bit (core/random_utils.py:20-87): per trajectory one
``SeedSequence([seed, traj, "TRAJ"])`` stream for initialize/step_through and one fresh
``SeedSequence([seed, traj, timestep, "SAMP"])`` stream per measurement copy.  A jump test consumes one double
(``rng.random()``) and, if it jumps, one more (``rng.choice`` draws exactly one uniform and selects with
``searchsorted(cdf, u, side="right")``), so a table of consecutive doubles per stream is an exact replay.

``philox`` mode fills the same tables on the device with the hand-written Philox4x32-10 kernel (``yb_philox_uniform``): key =
run seed, counter = (jump-test index, stream tag, global trajectory index), so row b depends only on ``traj_indices[b]`` and
results do not depend on how trajectories are batched or sharded.
"""
from __future__ import annotations

import numpy as np
import torch

STREAM_TRAJECTORY = 0x5452414A  # "TRAJ"
STREAM_SAMPLE = 0x53414D50      # "SAMP"
STREAM_DISORDER = 0x4449534F    # "DISO"
STREAM_SHOTS = 0x53484F54       # "SHOT" (ours: the reference samples shots from an unseeded Generator)


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


def philox_tables(kernels, seed: int | None, traj_indices, n_times: int, order: int, device, sample_offset: int = 0):
    """Device-side tables from the Philox4x32-10 kernel (csrc/yb_philox.cuh, ``yb_philox_uniform``): stream = (seed, global
    trajectory index, tag); the trajectory stream uses block c for jump test c, the sample stream uses block j + offset for
    the measurement copy of time step j -- the counter-based mirror of core/random_utils.py:20-87."""
    nb = len(traj_indices)
    base = int(seed) if seed is not None else int(np.random.SeedSequence().entropy % (2 ** 64))
    idx = torch.as_tensor(np.asarray(traj_indices, dtype=np.int64), device=device)
    traj = torch.empty((nb, 2 * n_times), dtype=torch.float64, device=device)
    kernels.philox_uniform(base, idx, STREAM_TRAJECTORY, n_times, 0, traj)
    samp = None
    if order == 2:
        flat = torch.empty((nb, 2 * n_times), dtype=torch.float64, device=device)
        kernels.philox_uniform(base, idx, STREAM_SAMPLE, n_times, sample_offset, flat)
        samp = flat.view(nb, n_times, 2)
    return traj, samp


def philox_source(seed: int | None, traj_indices, device, sample_offset: int = 0):
    """(seed, global trajectory indices int64 [B] on the device, sample offset): what ``BatchedTJM.philox`` takes to draw the uniforms
    of every jump test inside ``yb_jump_test`` -- the same numbers ``philox_tables`` would hold, without the tables."""
    base = int(seed) if seed is not None else int(np.random.SeedSequence().entropy % (2 ** 64))
    return base, torch.as_tensor(np.asarray(traj_indices, dtype=np.int64), device=device), int(sample_offset)
