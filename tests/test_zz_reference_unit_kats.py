"""Known answers of the reference's own UNIT tests for the functions on the hot path, asked of the batched engine.

Each test restates one test of the reference (cited) on `BatchedTJM`: once on the numpy kernel twin (`-m "not gpu"`) and once through
the C ABI (`-m gpu`).  Fixtures pin whole runs; these pin the single functions where the reference itself states an exact answer:

  stochastic_process   tests/core/methods/test_stochastic_process.py:264-296   (PDF [0.2, 0.8], [0.5, 0.5] independent of max_bond_dim)
  apply_dissipation    tests/core/methods/test_dissipation.py:25-94            (centre back at site 0), plus the exact norm it leaves
  tdvp                 tests/core/methods/tdvp/test_integrators.py:74-105,155-183 (2-site / 1-site step == expm(-i dt H) psi0, centre 0)
  jump probability     tests/analog/test_analog_tjm.py:250-320                 (dp of the first jump test against the dense MCWF value)
  seeded rerun         tests/core/test_random_utils.py:73-100                  (same seed => identical trajectories, atol 0)

The file sorts last on purpose: it re-asks questions the fixtures already answer, so it must never stop the suite before them.
The `-m gpu` twins were written after the round's GPU budget was spent; like config 5's parity test before its first run on a B200
they are opt-in (`YB_RUN_UNIT_KATS=1`) until they have been seen green on the device once.
"""
import os

import numpy as np
import pytest
import scipy.linalg
import torch

from oracle.np_kernels import NumpyKernels
from yaqs_b200 import AnalogSimParams, Hamiltonian, MPO, NoiseModel, Observable, Simulator, State, Z
from yaqs_b200.containers import MPS
from yaqs_b200.engine import BatchedTJM, EngineParams


@pytest.fixture(scope="module")
def cuda_kernels():
    if os.environ.get("YB_RUN_UNIT_KATS") != "1":
        pytest.skip("first run on a B200 pending: set YB_RUN_UNIT_KATS=1")
    from yaqs_b200.kernels import CudaKernels
    return CudaKernels("cuda:0")


def _engine(kernels, length, noise, init, batch=2, chi=8, order=1, dt=0.1, **kw):
    ep = EngineParams(dt=dt, n_steps=1, order=order, max_bond_dim=chi, svd_threshold=kw.pop("threshold", 1e-12), krylov_tol=kw.pop("krylov_tol", 1e-12),
                      tdvp_mode=kw.pop("tdvp_mode", "2site"))
    mpo = kw.pop("mpo", None) or [np.eye(2, dtype=np.complex128).reshape(2, 2, 1, 1) for _ in range(length)]
    procs = NoiseModel(noise).processes if noise else None
    obs = [(np.diag([1.0, -1.0]).astype(np.complex128), [0])]
    return BatchedTJM(kernels, length, batch, mpo, ep, procs, obs, init)


def _dense(tensors):
    v = np.ones((1, 1), dtype=np.complex128)
    for t in tensors:
        v = np.einsum("xl,slr->xsr", v, t).reshape(-1, t.shape[2])
    return v.reshape(-1)


def _forced_jump_pdf(kernels, noise, chi):
    """One jump test with u1 = 0 (dp > 0 => every trajectory jumps): the pdf the selection kernel built."""
    eng = _engine(kernels, 2, noise, MPS(2, state="zeros").tensors, chi=chi, dt=1.0)
    eng.apply_dissipation(eng.phi, 1.0)                  # dp must be positive for `u1 < dp`; the pdf is evaluated on this state
    b = eng.B
    u1 = torch.zeros(b, dtype=torch.float64, device=eng.dev)
    u2 = torch.full((b,), 0.3, dtype=torch.float64, device=eng.dev)
    log = []
    eng.stochastic_process(eng.phi, 1.0, u1, u2, log=log)
    assert log[0]["jumped"].all()
    assert int(eng.status.cpu().abs().sum()) == 0
    return log[0]["pdf"], log[0]["choice"]


# ------------------------------------------------------------------------------------------------------------------------------------
# stochastic_process: exact PDFs (test_stochastic_process.py:264-296)
# ------------------------------------------------------------------------------------------------------------------------------------
def _pdf_one_to_four(kernels):
    # product |00>: ||X|0>||^2 = 1, ||(2 I)|00>||^2 = 4 at equal rates -> [0.2, 0.8].  The dissipators of both processes are multiples
    # of the identity on |00> (X^dagger X = 1, (2I)^dagger (2I) = 4), so the state the pdf sees is |00> up to its norm, as in the reference
    noise = [{"name": "pauli_x", "sites": [0], "strength": 1.0},
             {"name": "scaled_i", "sites": [0, 1], "strength": 1.0, "matrix": 2.0 * np.eye(4, dtype=np.complex128)}]
    pdf, choice = _forced_jump_pdf(kernels, noise, chi=4)
    for row in pdf:
        np.testing.assert_allclose(row, [0.2, 0.8], atol=1e-10)
    assert (choice == 1).all()                           # u2 = 0.3 falls into [0.2, 1): numpy's choice rule picks the second process


def _pdf_independent_of_cap(kernels):
    # XX|00> = |11> and the Bell channel maps |00> to (|00> + |11>)/sqrt 2: both norms 1, but only the second needs chi = 2.  A pdf
    # taken after a truncated split at max_bond_dim = 1 would be [2/3, 1/3] (test_stochastic_process.py:277-296)
    x = np.array([[0, 1], [1, 0]], dtype=np.complex128)
    bell = np.zeros((4, 4), dtype=np.complex128)
    bell[:, 0] = [1 / np.sqrt(2), 0, 0, 1 / np.sqrt(2)]
    bell[:, 1] = [0, 1, 0, 0]
    bell[:, 2] = [0, 0, 1, 0]
    bell[:, 3] = [1 / np.sqrt(2), 0, 0, -1 / np.sqrt(2)]
    noise = [{"name": "xx", "sites": [0, 1], "strength": 1.0, "matrix": np.kron(x, x)},
             {"name": "bell", "sites": [0, 1], "strength": 1.0, "matrix": bell}]
    pdf, _ = _forced_jump_pdf(kernels, noise, chi=1)
    for row in pdf:
        np.testing.assert_allclose(row, [0.5, 0.5], atol=1e-10)


def test_pdf_weights_follow_operator_norms():
    _pdf_one_to_four(NumpyKernels())


def test_pdf_is_independent_of_max_bond_dim():
    _pdf_independent_of_cap(NumpyKernels())


@pytest.mark.gpu
def test_cuda_pdf_weights_follow_operator_norms(cuda_kernels):
    _pdf_one_to_four(cuda_kernels)


@pytest.mark.gpu
def test_cuda_pdf_is_independent_of_max_bond_dim(cuda_kernels):
    _pdf_independent_of_cap(cuda_kernels)


# ------------------------------------------------------------------------------------------------------------------------------------
# apply_dissipation: centre back at site 0 (test_dissipation.py:25-94) and the exact non-unitary factor
# ------------------------------------------------------------------------------------------------------------------------------------
def _dissipation_product_state(kernels, two_site):
    length, dt, gamma = 3, 0.1, 0.1
    rng = np.random.default_rng(11)
    vecs = [rng.random(2) + 0.2 for _ in range(length)]
    vecs = [v / np.linalg.norm(v) for v in vecs]
    init = MPS(length, tensors=[v.astype(np.complex128).reshape(2, 1, 1) for v in vecs]).tensors
    if two_site:
        noise = [{"name": n, "sites": [i, i + 1], "strength": gamma} for i in range(length - 1) for n in ("crosstalk_xx", "crosstalk_yy")]
    else:
        noise = [{"name": n, "sites": [i], "strength": gamma} for i in range(length) for n in ("lowering", "pauli_z")]
    eng = _engine(kernels, length, noise, init, chi=4)
    eng.apply_dissipation(eng.phi, dt)
    psi0 = _dense(init)
    # the exact image: exp(-dt/2 sum_k gamma_k L_k^dagger L_k) psi0 (dissipation.py:50-183; every L^dagger L here is diagonal or the identity)
    lower = np.array([[0, 1], [0, 0]], dtype=np.complex128)
    ident = np.eye(2, dtype=np.complex128)
    gen = np.zeros((8, 8), dtype=np.complex128)
    if two_site:
        gen += 2 * (length - 1) * gamma * np.eye(8)      # XX and YY are unitary: L^dagger L = 1 per process
    else:
        for i in range(length):
            ops = [ident] * length
            ops[i] = lower.conj().T @ lower
            gen += gamma * np.kron(np.kron(ops[0], ops[1]), ops[2]) + gamma * np.eye(8)
    want = scipy.linalg.expm(-0.5 * dt * gen) @ psi0
    for b in range(eng.B):
        got = eng.phi.download(b)
        assert np.max(np.abs(_dense(got) - want)) < 1e-12
        for t in got[1:]:                                # centre at site 0: every other site is a right isometry
            m = t.transpose(1, 0, 2).reshape(t.shape[1], -1)
            assert np.max(np.abs(m @ m.conj().T - np.eye(t.shape[1]))) < 1e-12
    assert int(eng.status.cpu().abs().sum()) == 0


@pytest.mark.parametrize("two_site", [False, True])
def test_dissipation_leaves_centre_at_zero_and_the_exact_factor(two_site):
    _dissipation_product_state(NumpyKernels(), two_site)


@pytest.mark.gpu
@pytest.mark.parametrize("two_site", [False, True])
def test_cuda_dissipation_leaves_centre_at_zero_and_the_exact_factor(cuda_kernels, two_site):
    _dissipation_product_state(cuda_kernels, two_site)


# ------------------------------------------------------------------------------------------------------------------------------------
# tdvp: one step equals the dense propagator at full bond dimension (test_integrators.py:74-105, 155-183)
# ------------------------------------------------------------------------------------------------------------------------------------
def _tdvp_exact(kernels, mode):
    length, dt = 5, 0.1
    mpo = MPO.ising(length, 1.0, 0.5)
    init = MPS(length, state="zeros")
    if mode == "1site":
        init.pad_bond_dimension(4)                       # 1TDVP cannot grow bonds: give it the full ones (test_integrators.py:155-183 pads too)
    eng = _engine(kernels, length, None, init.tensors, chi=4, mpo=[np.asarray(w) for w in mpo.tensors], tdvp_mode=mode, dt=dt)
    eng.tdvp(eng.phi)
    want = scipy.linalg.expm(-1j * dt * mpo.to_matrix()) @ _dense(init.tensors)
    tol = 1e-9 if mode == "2site" else 2e-3             # 1TDVP carries its splitting error (the reference's test only checks the centre)
    for b in range(eng.B):
        got = eng.phi.download(b)
        assert np.max(np.abs(_dense(got) - want)) < tol, float(np.max(np.abs(_dense(got) - want)))
        for t in got[1:]:
            m = t.transpose(1, 0, 2).reshape(t.shape[1], -1)
            assert np.max(np.abs(m @ m.conj().T - np.eye(t.shape[1]))) < 1e-10
    assert int(eng.status.cpu().abs().sum()) == 0


@pytest.mark.parametrize("mode", ["2site", "1site"])
def test_tdvp_step_equals_dense_propagator(mode):
    _tdvp_exact(NumpyKernels(), mode)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["2site", "1site"])
def test_cuda_tdvp_step_equals_dense_propagator(cuda_kernels, mode):
    _tdvp_exact(cuda_kernels, mode)


# ------------------------------------------------------------------------------------------------------------------------------------
# jump probability of the first test against the dense Monte-Carlo wave-function value (test_analog_tjm.py:250-320)
# ------------------------------------------------------------------------------------------------------------------------------------
def _first_jump_probability(kernels):
    length, dt, gamma = 4, 0.05, 0.3
    mpo = MPO.ising(length, 1.0, 0.7)
    init = MPS(length, state="x+")
    noise = [{"name": "lowering", "sites": [i], "strength": gamma} for i in range(length)]
    eng = _engine(kernels, length, noise, init.tensors, chi=4, mpo=[np.asarray(w) for w in mpo.tensors], dt=dt)
    eng.tdvp(eng.phi)                                    # order 1: U(dt), then D(dt), then the test (analog_tjm.py:369-454)
    eng.apply_dissipation(eng.phi, dt)
    b = eng.B
    log = []
    eng.stochastic_process(eng.phi, dt, torch.ones(b, dtype=torch.float64, device=eng.dev), torch.zeros(b, dtype=torch.float64, device=eng.dev), log=log)
    lower = np.array([[0, 1], [0, 0]], dtype=np.complex128)
    gen = np.zeros((16, 16), dtype=np.complex128)
    for i in range(length):
        ops = [np.eye(2, dtype=np.complex128)] * length
        ops[i] = lower.conj().T @ lower
        full = ops[0]
        for o in ops[1:]:
            full = np.kron(full, o)
        gen += gamma * full
    psi = scipy.linalg.expm(-0.5 * dt * gen) @ scipy.linalg.expm(-1j * dt * mpo.to_matrix()) @ _dense(init.tensors)
    want = 1.0 - float(np.vdot(psi, psi).real)
    assert not log[0]["jumped"].any()                    # u1 = 1 never jumps
    np.testing.assert_allclose(log[0]["dp"], want, atol=1e-10)


def test_first_jump_probability_matches_dense_mcwf():
    _first_jump_probability(NumpyKernels())


@pytest.mark.gpu
def test_cuda_first_jump_probability_matches_dense_mcwf(cuda_kernels):
    _first_jump_probability(cuda_kernels)


# ------------------------------------------------------------------------------------------------------------------------------------
# seeded rerun: identical trajectories, atol 0 (tests/core/test_random_utils.py:73-100)
# ------------------------------------------------------------------------------------------------------------------------------------
def _seeded_rerun(kernels, **sim_kw):
    length = 5

    def once():
        sp = AnalogSimParams(observables=[Observable(Z(), s) for s in range(length)], elapsed_time=0.4, dt=0.1, num_traj=6, max_bond_dim=4,
                             order=2, sample_timesteps=True, random_seed=7)
        noise = NoiseModel([{"name": n, "sites": [i], "strength": 0.3} for i in range(length) for n in ("lowering", "pauli_z")])
        return Simulator(show_progress=False, kernels=kernels, **sim_kw).run(State(length, initial="x+"), Hamiltonian.ising(length, 1.0, 0.5), sp, noise)

    a, b = once(), once()
    for x, y in zip(a.trajectories, b.trajectories):
        assert np.array_equal(np.asarray(x), np.asarray(y))
    assert np.array_equal(a.max_bond, b.max_bond) and np.array_equal(a.total_bond, b.total_bond)


def test_seeded_rerun_is_identical():
    _seeded_rerun(NumpyKernels())


@pytest.mark.gpu
def test_cuda_seeded_rerun_is_identical(cuda_kernels):
    _seeded_rerun(cuda_kernels, device="cuda:0")
