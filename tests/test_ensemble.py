"""Ensemble-level parity (BASELINE.json north_star, third bar): "ensemble averages must agree with the reference within
3 sigma statistical error on the same config".

Fixtures tests/golden/ensemble_*.npz come from the UNMODIFIED reference (oracle/make_ensemble_golden.py): the Lindblad
master-equation solution (analog/lindblad.py) and the mean / standard error of the reference's own TJM ensemble.  Our
ensemble is driven by the device Philox generator (rng_mode="philox"), i.e. it is statistically INDEPENDENT of the
reference's numpy streams -- this is the test that does not rely on replaying the reference's uniforms.

Criterion per (observable, time) point: |ours - ref| <= 3 * sqrt(sem_ours^2 + sem_ref^2); with ~60 points a Gaussian
ensemble exceeds 3 sigma somewhere with probability ~15 %, so we require that at most 2 % of the points exceed 3 sigma, none
exceeds 4.5 sigma, and the chi^2 over all points is consistent with its degrees of freedom (5 sigma band).  Against the exact
Lindblad solution the Trotter error of the scheme, O(dt^2) per unit time, is added as a 5e-3 allowance.
"""
import os

import numpy as np
import pytest

from yaqs_b200 import AnalogSimParams, Hamiltonian, NoiseModel, Observable, Simulator, State, Z

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

CASES = {
    "ensemble_ising6": dict(model="ising", L=6, J=1.0, g=0.5, init="zeros", chi=8, thr=1e-6, dt=0.1, steps=10,
                            noise=[("lowering", 0.1), ("pauli_z", 0.1)]),
    "ensemble_heis6": dict(model="heisenberg", L=6, J=1.0, g=0.0, init="Neel", chi=8, thr=1e-6, dt=0.1, steps=8,
                           noise=[("lowering", 0.2)]),
}


def _run(name, ntraj, seed, **sim_kw):
    cfg = CASES[name]
    L = cfg["L"]
    ham = Hamiltonian.ising(L, cfg["J"], cfg["g"]) if cfg["model"] == "ising" else Hamiltonian.heisenberg(L, cfg["J"], cfg["J"], cfg["J"])
    sp = AnalogSimParams(observables=[Observable(Z(), s) for s in range(L)], elapsed_time=cfg["dt"] * cfg["steps"], dt=cfg["dt"], num_traj=ntraj,
                         max_bond_dim=cfg["chi"], svd_threshold=cfg["thr"], order=2, sample_timesteps=True, random_seed=seed)
    noise = NoiseModel([{"name": n, "sites": [i], "strength": g} for i in range(L) for (n, g) in cfg["noise"]])
    res = Simulator(show_progress=False, rng_mode="philox", **sim_kw).run(State(L, initial=cfg["init"]), ham, sp, noise)
    rows = np.stack([np.asarray(res.trajectories[k]) for k in range(L)], axis=1)           # [traj, obs, T]
    return rows


def _check(name, rows):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    n = rows.shape[0]
    mean, sem = rows.mean(axis=0), rows.std(axis=0, ddof=1) / np.sqrt(n)
    live = (sem > 1e-12) | (g["ref_sem"] > 1e-12)                      # t = 0 (and frozen points) carry no noise: compare exactly
    assert np.max(np.abs(mean[~live] - g["ref_mean"][~live])) < 1e-9
    # (i) against the reference's own TJM ensemble (same scheme, same time-step error)
    z = (mean - g["ref_mean"])[live] / np.sqrt(sem ** 2 + g["ref_sem"] ** 2)[live]
    assert np.mean(np.abs(z) > 3.0) <= 0.02, f"{name}: {np.sum(np.abs(z) > 3)} of {z.size} points beyond 3 sigma"
    assert np.max(np.abs(z)) < 4.5
    # points are correlated in time, so chi^2 is only bounded loosely: mean z^2 must be O(1)
    assert np.mean(z ** 2) < 2.5, f"{name}: mean z^2 = {np.mean(z ** 2):.2f}"
    # (ii) against the exact Lindblad solution, with the scheme's O(dt^2) bias as an allowance
    dev = np.abs(mean - g["exact"])
    assert np.all(dev <= 3.5 * np.maximum(sem, 1e-12) + 5e-3 + 2.0 * np.abs(g["ref_mean"] - g["exact"])), f"{name}: max dev {dev.max():.4f}"
    return float(np.max(np.abs(z))), float(dev.max())


@pytest.mark.parametrize("name", sorted(CASES))
def test_philox_ensemble_matches_reference_host_twin(name):
    """Host logic + Philox addressing with the numpy kernel twin (small ensemble)."""
    from oracle.np_kernels import NumpyKernels
    rows = _run(name, 64, 314159, kernels=NumpyKernels())
    _check(name, rows)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_philox_ensemble_matches_reference_gpu(name):
    rows = _run(name, 1024, 271828, device="cuda:0")
    zmax, dev = _check(name, rows)
    # sharper than the host test: with 1024 trajectories the standard error is ~0.01
    assert dev < 0.05
