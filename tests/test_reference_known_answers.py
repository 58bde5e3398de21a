"""Known-answer vectors of the reference's own simulator tests (tests/test_simulator.py:858-1226) through yaqs_b200's public API.

The expected arrays are read from tests/golden/known_answers_simulator.json (oracle/make_known_answers.py extracts them from the
reference's test file); the set-ups below are those tests' own: a noise-free Ising chain L = 4, J = 1, g = 0.1 from |0000>, two-site
correlators XX / YY / ZZ at the left boundary, the centre and the right boundary every dt = 0.1 up to T = 2 (analog, default order,
max_bond_dim 4), and the same correlators at the end of the 20-step Trotter circuit (digital).  Tolerances are the reference's
(atol 1e-3 analog, 2e-3 circuit).  CPU: host engine on the numpy twin; -m gpu: the CUDA path."""
import json
import os

import numpy as np
import pytest
import torch

from yaqs_b200 import XX, YY, ZZ, AnalogSimParams, DigitalSimParams, Hamiltonian, Observable, Simulator, State
from yaqs_b200.circuits import create_ising_circuit

KNOWN = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "known_answers_simulator.json")))
L, J, G = 4, 1, 0.1
SITES = {"test_two_site_correlator_left_boundary": [0, 1], "test_two_site_correlator_center": [L // 2, L // 2 + 1],
         "test_two_site_correlator_right_boundary": [L - 2, L - 1]}


def _simulator(backend: str) -> Simulator:
    if backend == "twin":
        from oracle.np_kernels import NumpyKernels
        return Simulator(show_progress=False, kernels=NumpyKernels())
    return Simulator(show_progress=False, device="cuda:0")


def _check_analog(name: str, backend: str) -> None:
    sites = SITES[name]
    sp = AnalogSimParams(observables=[Observable(XX(), sites), Observable(YY(), sites), Observable(ZZ(), sites)], elapsed_time=2.0, dt=0.1,
                         max_bond_dim=4, sample_timesteps=True)
    result = _simulator(backend).run(State(L, initial="zeros"), Hamiltonian.ising(L, J, G), sp)
    for row, key in enumerate(("xx", "yy", "zz")):
        np.testing.assert_allclose(result.expectation_values[row], np.array(KNOWN[name]["expected"][key]), atol=1e-3)


def _check_circuit(backend: str) -> None:
    name = "test_two_site_correlator_center_circuit"
    sites = [L // 2, L // 2 + 1]
    sp = DigitalSimParams(observables=[Observable(XX(), sites), Observable(YY(), sites), Observable(ZZ(), sites)], max_bond_dim=4)
    result = _simulator(backend).run(State(L, initial="zeros"), create_ising_circuit(L=L, J=J, g=G, dt=0.1, timesteps=20), sp)
    for row, key in enumerate(("xx", "yy", "zz")):
        np.testing.assert_allclose(result.expectation_values[row], np.array(KNOWN[name]["expected"][key]), atol=2e-3)


@pytest.mark.parametrize("name", sorted(SITES))
def test_two_site_correlators_match_the_reference_known_answers(name):
    _check_analog(name, "twin")


def test_circuit_correlators_match_the_reference_known_answers():
    _check_circuit("twin")


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(SITES))
def test_cuda_two_site_correlators_match_the_reference_known_answers(name):
    if not torch.cuda.is_available():
        pytest.skip("needs CUDA")
    _check_analog(name, "cuda")


@pytest.mark.gpu
def test_cuda_circuit_correlators_match_the_reference_known_answers():
    if not torch.cuda.is_available():
        pytest.skip("needs CUDA")
    _check_circuit("cuda")


# ------------------------------------------------------------------------------------------------------------------------------------
# The two seeded golden vectors the reference's tests hold for the DIGITAL path (BASELINE config 4's path)
# ------------------------------------------------------------------------------------------------------------------------------------
# tests/digital/test_digital_tjm.py:1759-1821: 3 qubits, k x [rzz(0.5) on (0, 1), (1, 2)] sampled after every repetition, bit-flip +
# pauli_y on every qubit and crosstalk_xx / yy on neighbours at 0.01, 100 trajectories, random_seed 7, |000> padded to chi = 2;
# the reference compares with rtol 0, atol 1e-12.
NOISY_DIGITAL_REFERENCE = np.array([[1.0, 0.94, 0.82, 0.7, 0.52, 0.5],
                                    [1.0, 0.96, 0.86, 0.72, 0.46, 0.5],
                                    [1.0, 1.0, 0.94, 0.88, 0.82, 0.76]])
# tests/test_simulator.py:512-556: Ising circuit L = 5, J = 1, g = 0.5, dt = 0.1, 10 Trotter steps + measure_all, lowering + pauli_z at
# 1e-3, 10 trajectories, max_bond_dim 4, seed 42.  The vector in the reference's file is held at atol 2e-4 ("can drift slightly across
# platforms"); the second row is what the UNMODIFIED reference returns for this set-up in this container (digital_tjm through
# oracle/make_golden.run_digital_case, 18 Oct 2026), which the engine must reproduce to 1e-10.
DIGITAL_OBSERVABLES_EXPECTED = np.array([0.6733214071546825, 0.8502664720526317, 0.8709639049732125, 0.8628627940961556, 0.6730350827430835])
DIGITAL_OBSERVABLES_REFERENCE_HERE = np.array([0.6733228481722422, 0.8502749034190632, 0.8709912511167633, 0.8628392971319183, 0.6731236844049613])


def _check_noisy_digital(backend: str) -> None:
    from yaqs_b200 import NoiseModel, Z
    from yaqs_b200.circuits import SAMPLE_LABEL, Circuit
    n, strength = 3, 0.01
    noise = NoiseModel([{"name": "pauli_x", "sites": [i], "strength": strength} for i in range(n)]
                       + [{"name": "crosstalk_xx", "sites": [i, i + 1], "strength": strength} for i in range(n - 1)]
                       + [{"name": "crosstalk_yy", "sites": [i, i + 1], "strength": strength} for i in range(n - 1)]
                       + [{"name": "pauli_y", "sites": [i], "strength": strength} for i in range(n)])
    qc = Circuit(n)
    for k in range(5):
        qc.rzz(0.5, 0, 1)
        qc.rzz(0.5, 1, 2)
        if k < 4:
            qc.barrier(label=SAMPLE_LABEL)
    sp = DigitalSimParams(observables=[Observable(Z(), i) for i in range(n)], sample_layers=True, num_mid_measurements=4, num_traj=100, random_seed=7)
    result = _simulator(backend).run(State(n, initial="zeros", pad=2), qc, sp, noise)
    got = np.array([np.real(result.expectation_values[i][:6]) for i in range(n)])
    np.testing.assert_allclose(got, NOISY_DIGITAL_REFERENCE, rtol=0, atol=1e-12)


def _check_digital_observables(backend: str) -> None:
    from yaqs_b200 import NoiseModel, Z
    n = 5
    circuit = create_ising_circuit(L=n, J=1, g=0.5, dt=0.1, timesteps=10)
    circuit.measure_all()
    sp = DigitalSimParams(observables=[Observable(Z(), s) for s in range(n)], num_traj=10, max_bond_dim=4, krylov_tol=1e-12, random_seed=42)
    noise = NoiseModel([{"name": name, "sites": [i], "strength": 1e-3} for i in range(n) for name in ["lowering", "pauli_z"]])
    result = _simulator(backend).run(State(n, initial="zeros"), circuit, sp, noise)
    got = np.array([float(np.real(result.expectation_values[i][0])) for i in range(n)])
    assert all(len(t) == sp.num_traj for t in result.trajectories) and all(len(e) == 1 for e in result.expectation_values)
    np.testing.assert_allclose(got, DIGITAL_OBSERVABLES_EXPECTED, rtol=0, atol=2e-4)
    np.testing.assert_allclose(got, DIGITAL_OBSERVABLES_REFERENCE_HERE, rtol=0, atol=1e-10)


def test_noisy_digital_tjm_matches_the_reference_seeded_vector():
    _check_noisy_digital("twin")


def test_digital_observables_match_the_reference_seeded_vector():
    _check_digital_observables("twin")


_UNSEEN = pytest.mark.skipif(os.environ.get("YB_RUN_UNIT_KATS") != "1", reason="written after round 2's GPU budget was spent: opt-in until seen green on a B200")


@pytest.mark.gpu
@_UNSEEN
def test_cuda_noisy_digital_tjm_matches_the_reference_seeded_vector():
    _check_noisy_digital("cuda")


@pytest.mark.gpu
@_UNSEEN
def test_cuda_digital_observables_match_the_reference_seeded_vector():
    _check_digital_observables("cuda")
