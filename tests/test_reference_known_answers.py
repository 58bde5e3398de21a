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
