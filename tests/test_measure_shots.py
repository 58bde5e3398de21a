"""Projective shot sampling (MPS.measure_single_shot / measure_shots, mps.py:1282-1417) on the batched engine.

Fixture tests/golden/shots_random6.npz holds outcomes of the UNMODIFIED reference for two seeded random MPS with the uniforms
its numpy Generator consumed (oracle/make_shots_golden.py); replaying the uniforms must reproduce every outcome exactly."""
import os

import numpy as np
import pytest
import torch

from oracle import tjm_oracle as O
from yaqs_b200.engine import BatchedTJM, EngineParams

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _engine(kernels):
    g = np.load(os.path.join(GOLDEN, "shots_random6.npz"))
    length, chi = int(g["length"]), int(g["chi"])
    states = [O.random_mps(length, chi, int(s)).tensors for s in g["seeds"]]
    mpo = [np.eye(2, dtype=np.complex128).reshape(2, 2, 1, 1) for _ in range(length)]
    ep = EngineParams(dt=1.0, n_steps=1, order=1, max_bond_dim=chi, svd_threshold=1e-12, digital=True)
    eng = BatchedTJM(kernels, length, len(states), mpo, ep, None, [], states[0])
    eng.phi.load(states)
    return eng, g


def _check(eng, g):
    u = torch.from_numpy(g["u"]).to(eng.dev)
    before = [t.copy() for t in eng.phi.download(1)]
    got = eng.measure_shots(eng.phi, u).cpu().numpy()
    assert np.array_equal(got, g["outcomes"])
    for a, b in zip(before, eng.phi.download(1)):            # sampling leaves the state untouched
        assert np.array_equal(a, b)


def test_measure_shots_replays_reference_outcomes_host_twin():
    from oracle.np_kernels import NumpyKernels
    _check(*_engine(NumpyKernels()))


@pytest.mark.gpu
def test_measure_shots_replays_reference_outcomes_gpu():
    from yaqs_b200.kernels import CudaKernels
    _check(*_engine(CudaKernels("cuda:0")))


def test_shot_frequencies_follow_born_rule_host_twin():
    """Statistical check with fresh uniforms: frequencies of 4000 shots against |<bits|psi>|^2 of the dense state."""
    from oracle.np_kernels import NumpyKernels
    eng, g = _engine(NumpyKernels())
    length = int(g["length"])
    nshots = 2000
    u = torch.from_numpy(np.random.default_rng(5).random((2, nshots, length)))
    got = eng.measure_shots(eng.phi, u).cpu().numpy()
    for row, seed in enumerate(g["seeds"]):
        tens = O.random_mps(length, int(g["chi"]), int(seed)).tensors
        psi = tens[0][:, 0, :]
        for t in tens[1:]:
            psi = np.einsum("...a,sab->...sb", psi, t)
        psi = psi[..., 0]                                     # psi[s0, s1, ...]
        prob = np.abs(psi) ** 2
        prob /= prob.sum()
        flat = np.array([prob[tuple((k >> i) & 1 for i in range(length))] for k in range(2 ** length)])
        counts = np.bincount(got[row], minlength=2 ** length)
        big = flat * nshots >= 5
        chi2 = float((((counts - flat * nshots) ** 2) / np.maximum(flat * nshots, 1e-30))[big].sum())
        dof = int(big.sum()) - 1
        assert chi2 < dof + 5 * np.sqrt(2 * dof), (chi2, dof)


def test_digital_shots_through_simulator_host_twin():
    """DigitalSimParams(shots=...) through Simulator.run (digital_tjm.py:728-747, simulator.py:1001-1050): a noise-free circuit
    samples the whole budget from the one final state (counts follow the dense Born probabilities); with noise and observables
    the budget is spread over the trajectory ensemble; shots-only with noise runs one trajectory per shot."""
    from oracle.np_kernels import NumpyKernels
    from yaqs_b200 import DigitalSimParams, NoiseModel, Observable, Simulator, State, Z
    from yaqs_b200.circuits import create_ising_circuit
    L, nshots = 4, 1500
    circ = create_ising_circuit(L, 1.0, 0.7, 0.2, 2)
    res = Simulator(show_progress=False, kernels=NumpyKernels()).run(State(L, initial="zeros"), circ,
                                                                    DigitalSimParams(shots=nshots, max_bond_dim=None, svd_threshold=1e-12, random_seed=7))
    assert len(res.measurements) == 1 and sum(res.counts.values()) == nshots and list(res.counts) == sorted(res.counts)
    x, z = np.array([[0, 1], [1, 0]], dtype=complex), np.diag([1.0 + 0j, -1.0])

    def op_on(mats):
        out = np.eye(1)
        for s in range(L):
            out = np.kron(out, mats.get(s, np.eye(2)))                         # site 0 = most significant (MPS order)
        return out
    psi = np.zeros(2 ** L, dtype=complex)
    psi[0] = 1
    a, b = -2 * 0.2 * 0.7, -2 * 0.2 * 1.0
    rx = np.cos(a / 2) * np.eye(2) - 1j * np.sin(a / 2) * x
    for _ in range(2):
        for s in range(L):
            psi = op_on({s: rx}) @ psi
        for s in range(L - 1):
            psi = (np.cos(b / 2) * np.eye(2 ** L) - 1j * np.sin(b / 2) * op_on({s: z, s + 1: z})) @ psi
    prob = np.abs(psi) ** 2                                                     # index: site 0 most significant
    want = np.array([prob[int("".join(str((k >> i) & 1) for i in range(L)), 2)] for k in range(2 ** L)])     # key k: bit i = site i
    got = np.array([res.counts.get(k, 0) for k in range(2 ** L)])
    big = want * nshots >= 5
    chi2 = float((((got - want * nshots) ** 2) / np.maximum(want * nshots, 1e-30))[big].sum())
    dof = int(big.sum()) - 1
    assert chi2 < dof + 5 * np.sqrt(2 * dof), (chi2, dof)
    assert got[~big].sum() <= 5 * max(1.0, float((want[~big] * nshots).sum()))
    # same seed => same counts (the reference's shots are not reproducible; ours are)
    res_b = Simulator(show_progress=False, kernels=NumpyKernels()).run(State(L, initial="zeros"), circ,
                                                                      DigitalSimParams(shots=nshots, max_bond_dim=None, svd_threshold=1e-12, random_seed=7))
    assert res_b.counts == res.counts
    # noisy + observables: 10 shots over 4 trajectories -> 3, 3, 2, 2
    noise = NoiseModel([{"name": "lowering", "sites": [i], "strength": 0.05} for i in range(L)])
    sp = DigitalSimParams(observables=[Observable(Z(), 0)], num_traj=4, shots=10, max_bond_dim=8, random_seed=3)
    res = Simulator(show_progress=False, kernels=NumpyKernels()).run(State(L, initial="zeros"), circ, sp, noise)
    assert [sum(m.values()) for m in res.measurements] == [3, 3, 2, 2] and sum(res.counts.values()) == 10
    assert len(res.trajectories[0]) == 4
    # noisy, shots only: one trajectory per shot, num_traj ignored
    sp = DigitalSimParams(num_traj=2, shots=6, max_bond_dim=8, random_seed=3)
    res = Simulator(show_progress=False, kernels=NumpyKernels()).run(State(L, initial="zeros"), circ, sp, noise)
    assert len(res.measurements) == 6 and all(sum(m.values()) == 1 for m in res.measurements) and sum(res.counts.values()) == 6
