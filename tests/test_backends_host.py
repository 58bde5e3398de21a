"""The reference's per-trajectory backend boundary: analog_tjm_1 / analog_tjm_2 / digital_tjm keep their argument
tuple and return shapes as batches of one trajectory; run_batch is the same call for many trajectory indices.  Checked against the
fixtures of the unmodified reference with the numpy kernel twin (the -m gpu twin of this file runs the CUDA library)."""
import os

import numpy as np
import pytest

from oracle.golden_cases import CASES, DIGITAL_CASES
from yaqs_b200 import AnalogSimParams, DigitalSimParams, Hamiltonian, MPS, NoiseModel, Observable, State, Z, ZZ
from yaqs_b200 import backends
from yaqs_b200.circuits import create_ising_circuit
from yaqs_b200.rng import STREAM_TRAJECTORY

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _analog_inputs(name):
    cfg = CASES[name]
    L = cfg["L"]
    obs = []
    for tok in cfg["obs"].split("+"):
        obs += [Observable(Z(), s) for s in range(L)] if tok == "z" else [Observable(ZZ(), [s, s + 1]) for s in range(L - 1)]
    noise = NoiseModel([{"name": n, "sites": [i], "strength": g} for i in range(L) for (n, g) in cfg["noise"]])
    sp = AnalogSimParams(observables=obs, elapsed_time=cfg["dt"] * cfg["steps"], dt=cfg["dt"], num_traj=cfg["ntraj"], max_bond_dim=cfg["chi"],
                         svd_threshold=cfg["thr"], krylov_tol=1e-4, order=cfg["order"], sample_timesteps=cfg["sample_timesteps"],
                         random_seed=cfg["seed"])
    ham = Hamiltonian.ising(L, cfg["J"], cfg["g"]) if cfg["model"] == "ising" else Hamiltonian.heisenberg(L, cfg["J"], cfg["J"], cfg["J"])
    return cfg, State(L, initial=cfg["init"]), noise, sp, ham, np.load(os.path.join(GOLDEN, name + ".npz"))


def _check_backends(kernels):
    backends.set_default_kernels(kernels)
    try:
        # order 2: one trajectory through the reference's signature, with the standalone seed and with an external Generator
        cfg, state, noise, sp, ham, g = _analog_inputs("ising10_c1_order2")
        for t in (0, 3):
            res, diag, fin = backends.analog_tjm_2((t, state.mps, noise, sp, ham.mpo))
            assert res.shape == g["results"][t].shape and diag.shape == (3, res.shape[1]) and fin is None
            assert np.max(np.abs(res - g["results"][t])) < 1e-10 and np.array_equal(diag, g["diagnostics"][t])
        rng = np.random.default_rng(np.random.SeedSequence([cfg["seed"], 3, STREAM_TRAJECTORY]))
        res_b, _, _ = backends.analog_tjm_2((3, state, noise, sp, ham), rng=rng, copy_initial_state=False)
        assert np.max(np.abs(res_b - res)) < 1e-12
        with pytest.raises(ValueError, match="order"):
            backends.analog_tjm_1((0, state, noise, sp, ham))
        with pytest.raises(ValueError, match="outside the time grid"):
            backends.analog_tjm_2((0, state, noise, sp, ham), sample_at=[len(sp.times)])
        # the batched entry: several trajectory indices in lock step, rows in sorted-observable order
        res_all, diag_all, status, final = backends.run_batch([5, 1, 2], state, noise, sp, ham)
        assert not status.any() and final is None
        for row, t in enumerate([5, 1, 2]):
            assert np.max(np.abs(res_all[row] - g["results"][t])) < 1e-10 and np.array_equal(diag_all[row], g["diagnostics"][t])
        # order 1, Heisenberg, Z and ZZ
        cfg, state, noise, sp, ham, g = _analog_inputs("ising10_c1_order1")
        res, diag, _ = backends.analog_tjm_1((2, state, noise, sp, ham))
        assert np.max(np.abs(res - g["results"][2])) < 1e-10 and np.array_equal(diag, g["diagnostics"][2])
        # noise-free with get_state: the final MPS comes back normalised
        sp2 = AnalogSimParams(observables=[Observable(Z(), 0)], elapsed_time=0.2, dt=0.1, max_bond_dim=8, order=1, get_state=True)
        _, _, fin = backends.analog_tjm_1((0, State(6, initial="x+"), None, sp2, Hamiltonian.ising(6, 1.0, 0.5)))
        assert isinstance(fin, MPS) and fin.length == 6
        # digital
        cfg = DIGITAL_CASES["dig_ising6_nn_noisy"]
        L = cfg["L"]
        circ = create_ising_circuit(L, cfg["J"], cfg["g"], cfg["dt"], cfg["timesteps"], sample_every_step=True)
        spd = DigitalSimParams(observables=[Observable(Z(), s) for s in range(L)], num_traj=cfg["ntraj"], max_bond_dim=cfg["chi"],
                               svd_threshold=cfg["thr"], sample_layers=True, random_seed=cfg["seed"])
        noise = NoiseModel([{"name": n, "sites": [i], "strength": s_} for i in range(L) for (n, s_) in cfg["noise"]])
        gd = np.load(os.path.join(GOLDEN, "dig_ising6_nn_noisy.npz"))
        res, diag, counts, fin = backends.digital_tjm((4, State(L, initial="zeros"), noise, spd, circ))
        assert counts is None and fin is None
        assert np.max(np.abs(res - gd["results"][4])) < 1e-10 and np.array_equal(diag, gd["diagnostics"][4])
        spd2 = DigitalSimParams(shots=7, num_traj=3, max_bond_dim=cfg["chi"], random_seed=1)
        res, diag, counts, _ = backends.digital_tjm((0, State(L, initial="zeros"), None, spd2, circ))
        assert res is None and diag is None and sum(counts.values()) == 7
    finally:
        backends.set_default_kernels(None)


def test_backend_boundary_host_twin():
    from oracle.np_kernels import NumpyKernels
    _check_backends(NumpyKernels())


@pytest.mark.gpu
def test_backend_boundary_gpu():
    from yaqs_b200.kernels import CudaKernels
    _check_backends(CudaKernels("cuda:0"))


def test_backends_fail_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a machine without CUDA")
    backends.set_default_kernels(None)
    cfg, state, noise, sp, ham, _ = _analog_inputs("ising10_c1_order2")
    with pytest.raises(Exception):
        backends.analog_tjm_2((0, state, noise, sp, ham))


def _check_program_hand_off_through_backends(kernels, golden_dir):
    """The reference's program executor drives the two analog backends with ONE Generator and the hand-off keywords
    (simulator.py:339-402): segment 0 returns phi, segment 1 continues the Trotter chain with the global sample offset, segment 2
    (order 1) starts from the last physical copy.  Replayed here through backends.analog_tjm_2 / analog_tjm_1 for every trajectory of
    the fixture prog_handoff_o2 (unmodified reference)."""
    import os

    from oracle.golden_cases import PROGRAM_CASES
    from yaqs_b200 import MPS
    from yaqs_b200.rng import STREAM_TRAJECTORY
    backends.set_default_kernels(kernels)
    cfg = PROGRAM_CASES["prog_handoff_o2"]
    g = np.load(os.path.join(golden_dir, "prog_handoff_o2.npz"))
    L = cfg["L"]
    obs = [Observable(Z(), s) for s in range(L)]
    noise = NoiseModel([{"name": n, "sites": [i], "strength": s} for i in range(L) for (n, s) in cfg["noise"]])
    hams = {"h0": Hamiltonian.ising(L, 1.0, 0.8), "h1": Hamiltonian.ising(L, 0.5, 1.2)}
    sps = [AnalogSimParams(observables=obs, elapsed_time=seg["T"], dt=seg["dt"], max_bond_dim=seg["chi"], svd_threshold=seg["thr"],
                           order=seg["order"], sample_timesteps=True, get_state=True, random_seed=cfg["seed"]) for seg in cfg["segments"]]
    for t in range(cfg["ntraj"]):
        rng = np.random.default_rng(np.random.SeedSequence([cfg["seed"], t, STREAM_TRAJECTORY]))
        state = MPS(L, state=cfg["init"])
        r0, d0, state = backends.analog_tjm_2((t, state, noise, sps[0], hams["h0"]), copy_initial_state=False, rng=rng, sample_timestep_offset=0,
                                              return_trajectory_state=True, continue_trajectory=False)
        r1, d1, state = backends.analog_tjm_2((t, state, noise, sps[1], hams["h0"]), copy_initial_state=False, rng=rng, sample_timestep_offset=2,
                                              return_trajectory_state=False, continue_trajectory=True)
        r2, d2, state = backends.analog_tjm_1((t, state, noise, sps[2], hams["h1"]), copy_initial_state=False, rng=rng)
        for k, (r, d) in enumerate(((r0, d0), (r1, d1), (r2, d2))):
            assert np.max(np.abs(r - g[f"res_{k}"][t])) < 1e-10, (t, k)
            assert np.array_equal(d, g[f"diag_{k}"][t]), (t, k)
        # the shared Generator advanced by exactly the doubles the reference drew
        fresh = np.random.default_rng(np.random.SeedSequence([cfg["seed"], t, STREAM_TRAJECTORY]))
        fresh.random(int(g["draws"][t]))
        assert rng.random() == fresh.random()


def test_program_hand_off_through_backends_host_twin(golden_dir):
    from oracle.np_kernels import NumpyKernels
    try:
        _check_program_hand_off_through_backends(NumpyKernels(), golden_dir)
    finally:
        backends.set_default_kernels(None)


def _check_schmidt(kernels):
    """schmidt_spectrum at the backend boundary (missing #7): analog_tjm_1 returns a dtype=object array whose
    Schmidt rows hold the 500-vectors (analog_tjm.py:424-426, mps.py:644-678); fixture = the unmodified reference's own call
    (oracle/make_schmidt_golden.py).  analog_tjm_2 and Simulator.run fail the way the reference fails (float result buffers)."""
    from yaqs_b200 import Simulator
    g = np.load(os.path.join(GOLDEN, "schmidt_ising6_order1.npz"))
    L = int(g["L"])
    obs = [Observable(Z(), s) for s in range(L)] + [Observable("schmidt_spectrum", [s, s + 1]) for s in range(L - 1)]
    noise = NoiseModel([{"name": n, "sites": [i], "strength": float(g["gamma"])} for i in range(L) for n in ("lowering", "pauli_z")])
    sp = AnalogSimParams(observables=obs, elapsed_time=float(g["dt"]) * int(g["steps"]), dt=float(g["dt"]), num_traj=int(g["ntraj"]),
                         max_bond_dim=int(g["chi"]), svd_threshold=float(g["thr"]), order=1, sample_timesteps=True, random_seed=int(g["seed"]))
    names = [o.gate.name for o in sp.sorted_observables]
    assert names == [str(n) for n in g["names"]]
    ham = Hamiltonian.ising(L, float(g["J"]), float(g["g"]))
    backends.set_default_kernels(kernels)
    try:
        for t in range(int(g["ntraj"])):
            res, diag, _ = backends.analog_tjm_1((t, State(L, initial="x+"), noise, sp, ham))
            assert res.dtype == object and res.shape == (len(names), int(g["steps"]) + 1)
            for k, name in enumerate(names):
                for col in range(res.shape[1]):
                    if name == "schmidt_spectrum":
                        got, want = np.asarray(res[k, col], dtype=np.float64), g["spectra"][t, k, col]
                        assert got.shape == (500,) and np.array_equal(np.isnan(got), np.isnan(want))
                        assert np.nanmax(np.abs(got - want)) < 1e-10
                    else:
                        assert abs(float(res[k, col]) - g["scalars"][t, k, col]) < 1e-10
        sp2 = AnalogSimParams(observables=obs, elapsed_time=0.2, dt=0.1, num_traj=1, max_bond_dim=8, order=2, random_seed=1)
        with pytest.raises(ValueError, match="setting an array element with a sequence"):
            backends.analog_tjm_2((0, State(L, initial="x+"), noise, sp2, ham))
        with pytest.raises(ValueError, match="setting an array element with a sequence"):
            Simulator(show_progress=False, kernels=kernels).run(State(L, initial="x+"), ham, sp, noise)
    finally:
        backends.set_default_kernels(None)


def test_schmidt_spectrum_at_the_backend_boundary():
    from oracle.np_kernels import NumpyKernels
    _check_schmidt(NumpyKernels())


@pytest.mark.gpu
def test_schmidt_spectrum_at_the_backend_boundary_gpu():
    from yaqs_b200.kernels import CudaKernels
    _check_schmidt(CudaKernels("cuda:0"))
