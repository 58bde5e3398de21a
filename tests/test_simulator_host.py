"""-m "not gpu": the public Simulator surface (validation, noise sampling, observable ordering, Result assembly, trajectory
sharding over torch.distributed) with the numpy kernel twin injected in place of the CUDA library."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from oracle.np_kernels import NumpyKernels
from yaqs_b200 import AnalogSimParams, Hamiltonian, MPO, NoiseModel, Observable, Simulator, State, X, Z, ZZ
from yaqs_b200.simulator import shard_range

REF_GOLDEN_Z = [0.748947146695782, 0.8720515025769692, 0.8652609567462763, 0.8673233347433466, 0.6872036335377433]


def _golden_run(**kw):
    length = 5
    sp = AnalogSimParams(observables=[Observable(Z(), s) for s in range(length)], elapsed_time=1, dt=0.1, num_traj=10, max_bond_dim=4,
                         svd_threshold=1e-6, order=2, sample_timesteps=False, random_seed=42)
    noise = NoiseModel([{"name": n, "sites": [i], "strength": 0.1} for i in range(length) for n in ["lowering", "pauli_z"]])
    return Simulator(show_progress=False, kernels=NumpyKernels(), **kw).run(State(length, initial="zeros"), Hamiltonian.ising(length, J=1, g=0.5),
                                                                          sp, noise), sp


def test_reference_golden_through_simulator():
    """The reference's own seeded test (tests/test_simulator.py:159-207, atol there 2e-4) through our public API."""
    res, sp = _golden_run()
    got = np.array([res.expectation_values[i][0] for i in range(5)])
    assert np.max(np.abs(got - REF_GOLDEN_Z)) < 1e-10
    assert all(len(t) == sp.num_traj for t in res.trajectories)
    assert res.times.shape == (1,) and res.max_bond.shape == (1,)


def test_waves_do_not_change_results():
    a, _ = _golden_run()
    b, _ = _golden_run(batch_size=3)
    for x, y in zip(a.trajectories, b.trajectories):
        assert np.array_equal(x, y)


def test_user_order_and_two_site_observables():
    length = 4
    obs = [Observable(ZZ(), [1, 2]), Observable(X(), 3), Observable(Z(), 0), Observable("x", 0)]
    sp = AnalogSimParams(observables=obs, elapsed_time=0.2, dt=0.1, max_bond_dim=4, order=1)
    res = Simulator(kernels=NumpyKernels()).run(State(length, initial="x+"), Hamiltonian.ising(length, 1.0, 0.5), sp)
    assert len(res.trajectories[0]) == 1            # noise-free => a single trajectory (simulator.py:1549-1555)
    import scipy.linalg
    h = MPO.ising(length, 1.0, 0.5).to_matrix()
    psi0 = np.ones(2 ** length) / 4.0
    zz, x, z = np.kron(np.diag([1, -1]), np.diag([1, -1])), np.array([[0, 1], [1, 0]]), np.diag([1.0, -1.0])
    full = [np.kron(np.kron(np.eye(2), zz), np.eye(2)), np.kron(np.eye(8), x), np.kron(z, np.eye(8)), np.kron(x, np.eye(8))]
    for col, t in enumerate(sp.times):
        psi = scipy.linalg.expm(-1j * t * h) @ psi0
        for k, op in enumerate(full):
            assert abs(res.expectation_values[k][col] - np.vdot(psi, op @ psi).real) < 5e-5


def test_entropy_and_pvm_observables_through_simulator(golden_dir):
    """Meta-observables through the public API (Observable("entropy", [i, i+1]); any non-GateLibrary string is a PVM bitstring,
    simulation_parameters.py:388-400) against the fixtures of the unmodified reference."""
    from oracle.golden_cases import CASES
    cfg = CASES["ising6_entropy_order2"]
    L = cfg["L"]
    obs = [Observable(Z(), s) for s in range(L)] + [Observable("entropy", [s, s + 1]) for s in range(L - 1)]
    noise = NoiseModel([{"name": n, "sites": [i], "strength": g} for i in range(L) for (n, g) in cfg["noise"]])
    sp = AnalogSimParams(observables=obs, elapsed_time=cfg["dt"] * cfg["steps"], dt=cfg["dt"], num_traj=cfg["ntraj"], max_bond_dim=cfg["chi"],
                         svd_threshold=cfg["thr"], krylov_tol=1e-4, order=2, sample_timesteps=True, random_seed=cfg["seed"])
    res = Simulator(show_progress=False, kernels=NumpyKernels()).run(State(L, initial=cfg["init"]), Hamiltonian.ising(L, cfg["J"], cfg["g"]), sp, noise)
    g = np.load(os.path.join(golden_dir, "ising6_entropy_order2.npz"))
    idx = sp.observable_sorted_indices
    for user_i in range(len(obs)):
        assert np.max(np.abs(np.asarray(res.trajectories[user_i]) - g["results"][:, idx[user_i], :])) < 1e-10
    cfg = CASES["ising6_pvm_order2"]
    pv = [Observable("0" * L), Observable(("01" * L)[:L]), Observable(("110" * L)[:L])]
    sp = AnalogSimParams(observables=pv, elapsed_time=cfg["dt"] * cfg["steps"], dt=cfg["dt"], num_traj=cfg["ntraj"], max_bond_dim=cfg["chi"],
                         svd_threshold=cfg["thr"], krylov_tol=1e-4, order=2, sample_timesteps=True, random_seed=cfg["seed"])
    res = Simulator(show_progress=False, kernels=NumpyKernels()).run(State(L, initial=cfg["init"]), Hamiltonian.ising(L, cfg["J"], cfg["g"]), sp, noise)
    g = np.load(os.path.join(golden_dir, "ising6_pvm_order2.npz"))
    for k in range(3):
        assert np.max(np.abs(np.asarray(res.trajectories[k]) - g["results"][:, k, :])) < 1e-10
    with pytest.raises(AssertionError, match="mixed observable and projective-measurement"):
        AnalogSimParams(observables=[Observable(Z(), 0), Observable("0" * L)], elapsed_time=0.1, dt=0.1)
    assert Observable("schmidt_spectrum", [0, 1]).gate.name == "schmidt_spectrum"      # accepted; see test_backends_host for where it runs


def test_scheduled_jumps_through_simulator(golden_dir):
    """NoiseModel(processes, scheduled_jumps=[...]) with AnalogSimParams(order=1) through the public API against the fixture of
    the unmodified reference; the reference's validation errors (order 2, off-grid time, non-adjacent sites, digital runs)."""
    from oracle.golden_cases import CASES
    from yaqs_b200 import DigitalSimParams
    from yaqs_b200.circuits import create_ising_circuit
    cfg = CASES["ising6_scheduled_order1"]
    L = cfg["L"]
    obs = [Observable(Z(), s) for s in range(L)] + [Observable(ZZ(), [s, s + 1]) for s in range(L - 1)]
    procs = [{"name": n, "sites": [i], "strength": g} for i in range(L) for (n, g) in cfg["noise"]]
    sched = [{"time": t, "sites": list(st), "name": n} for (t, st, n) in cfg["scheduled"]]
    noise = NoiseModel(procs, scheduled_jumps=sched)
    kw = dict(elapsed_time=cfg["dt"] * cfg["steps"], dt=cfg["dt"], num_traj=cfg["ntraj"], max_bond_dim=cfg["chi"], svd_threshold=cfg["thr"],
              krylov_tol=1e-4, sample_timesteps=True, random_seed=cfg["seed"])
    sp = AnalogSimParams(observables=obs, order=1, **kw)
    res = Simulator(show_progress=False, kernels=NumpyKernels()).run(State(L, initial=cfg["init"]), Hamiltonian.ising(L, cfg["J"], cfg["g"]), sp, noise)
    g = np.load(os.path.join(golden_dir, "ising6_scheduled_order1.npz"))
    idx = sp.observable_sorted_indices
    for user_i in range(len(obs)):
        assert np.max(np.abs(np.asarray(res.trajectories[user_i]) - g["results"][:, idx[user_i], :])) < 1e-10
    sim, st, ham = Simulator(show_progress=False, kernels=NumpyKernels()), State(L, initial="zeros"), Hamiltonian.ising(L, 1.0, 0.7)
    with pytest.raises(ValueError, match="only supported for AnalogSimParams\\(order=1\\)"):
        sim.run(st, ham, AnalogSimParams(observables=obs, order=2, **kw), noise)
    with pytest.raises(ValueError, match="not on the simulation time grid"):
        sim.run(st, ham, sp, NoiseModel(procs, scheduled_jumps=[{"time": 0.25, "sites": [0], "name": "pauli_x"}]))
    with pytest.raises(ValueError, match="non-adjacent"):
        NoiseModel(procs, scheduled_jumps=[{"time": 0.2, "sites": [0, 2], "name": "crosstalk_xx"}])
    with pytest.raises(ValueError, match="annihilates the current state"):
        sim.run(st, ham, sp, NoiseModel(procs, scheduled_jumps=[{"time": 0.0, "sites": [0], "name": "lowering"}]))
    with pytest.raises(ValueError, match="scheduled_jumps are only supported for single-State analog"):
        sim.run(st, create_ising_circuit(L, 1.0, 0.7, 0.2, 1), DigitalSimParams(observables=[Observable(Z(), 0)], num_traj=2), noise)


def test_mpo_factories_match_dense():
    z, x, y, i2 = np.diag([1.0, -1.0]), np.array([[0, 1.0], [1, 0]]), np.array([[0, -1j], [1j, 0]]), np.eye(2)

    def emb(op, site, n):
        out = np.eye(1)
        for s in range(n):
            out = np.kron(out, op if s == site else i2)
        return out

    n = 4
    h = sum(-1.3 * emb(z, s, n) @ emb(z, s + 1, n) for s in range(n - 1)) + sum(-0.7 * emb(x, s, n) for s in range(n))
    assert np.allclose(MPO.ising(n, 1.3, 0.7).to_matrix(), h)
    hp = h - 1.3 * emb(z, n - 1, n) @ emb(z, 0, n)
    assert np.allclose(MPO.ising(n, 1.3, 0.7, bc="periodic").to_matrix(), hp)
    hh = sum(-(0.5 * emb(x, s, n) @ emb(x, s + 1, n) + 0.8 * emb(y, s, n) @ emb(y, s + 1, n) + 1.1 * emb(z, s, n) @ emb(z, s + 1, n))
             for s in range(n - 1)) + sum(-0.3 * emb(z, s, n) for s in range(n))
    assert np.allclose(MPO.heisenberg(n, 0.5, 0.8, 1.1, 0.3).to_matrix(), hh)


def test_validation_errors_match_reference():
    with pytest.raises(ValueError, match="integer multiple of dt"):
        AnalogSimParams(elapsed_time=0.25, dt=0.1)
    with pytest.raises(ValueError, match="Unknown noise operator"):
        NoiseModel([{"name": "bogus", "sites": [0], "strength": 0.1}])
    with pytest.raises(ValueError, match="nonnegative"):
        NoiseModel([{"name": "lowering", "sites": [0], "strength": -0.1}])
    with pytest.raises(ValueError, match="No output specified"):
        Simulator(kernels=NumpyKernels()).run(State(3), Hamiltonian.ising(3, 1, 1), AnalogSimParams())
    sp = AnalogSimParams(observables=[Observable(Z(), 0)], get_state=True)
    with pytest.raises(ValueError, match="Cannot return state in noisy"):
        Simulator(kernels=NumpyKernels()).run(State(3), Hamiltonian.ising(3, 1, 1), sp, NoiseModel([{"name": "lowering", "sites": [0], "strength": 0.1}]))


def test_product_without_cuda_fails_loudly():
    """No CPU fallback: without a CUDA device the default backend must raise, not degrade."""
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from yaqs_b200.kernels import KernelError
    sp = AnalogSimParams(observables=[Observable(Z(), 0)])
    with pytest.raises(KernelError):
        Simulator().run(State(3), Hamiltonian.ising(3, 1, 1), sp)


def test_shard_range_partitions():
    for n in (0, 1, 7, 4096):
        for w in (1, 2, 3, 8):
            ids = [i for r in range(w) for i in shard_range(n, r, w)]
            assert ids == list(range(n))


def _digital_shots_run():
    """Noisy digital run with observables and a shot budget spread over 5 trajectories (3 + 2 on two ranks)."""
    from yaqs_b200 import DigitalSimParams
    from yaqs_b200.circuits import create_ising_circuit
    L = 4
    noise = NoiseModel([{"name": "lowering", "sites": [i], "strength": 0.1} for i in range(L)])
    sp = DigitalSimParams(observables=[Observable(Z(), 1)], num_traj=5, shots=12, max_bond_dim=8, random_seed=11)
    return Simulator(show_progress=False, kernels=NumpyKernels()).run(State(L, initial="zeros"), create_ising_circuit(L, 1.0, 0.7, 0.2, 2), sp, noise)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        res, _ = _golden_run()
        dig = _digital_shots_run()
        if rank == 0:
            out.put((np.array([res.expectation_values[i][0] for i in range(5)]), np.stack([t[:, 0] for t in res.trajectories]),
                     dig.counts, [dict(m) for m in dig.measurements], np.asarray(dig.trajectories[0])))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharding_matches_single_process():
    """N > 1 path on CPU (gloo, world_size 2): trajectories are sharded by index range, sums combined by one all-reduce."""
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    means, traj, counts, measurements, dig_traj = q.get(timeout=600)
    for p in procs:
        p.join(timeout=600)
        assert p.exitcode == 0
    single, _ = _golden_run()
    assert np.max(np.abs(means - REF_GOLDEN_Z)) < 1e-10
    assert np.array_equal(traj, np.stack([t[:, 0] for t in single.trajectories]))
    # digital path with shots: per-trajectory counts are gathered from both ranks and do not depend on the sharding
    dig = _digital_shots_run()
    assert counts == dig.counts and measurements == [dict(m) for m in dig.measurements] and sum(counts.values()) == 12
    assert [sum(m.values()) for m in measurements] == [3, 3, 2, 2, 2]
    assert np.array_equal(dig_traj, np.asarray(dig.trajectories[0]))


def test_digital_simulator_matches_reference_fixture(golden_dir):
    """Simulator.run(State, Circuit, DigitalSimParams, NoiseModel): the digital TJM through the public API reproduces the
    unmodified reference's per-trajectory observables (fixture dig_ising6_nn_noisy) and a noise-free circuit matches the
    dense state-vector evolution of the same gate sequence."""
    from oracle.golden_cases import DIGITAL_CASES
    from yaqs_b200 import DigitalSimParams
    from yaqs_b200.circuits import create_ising_circuit
    cfg = DIGITAL_CASES["dig_ising6_nn_noisy"]
    L = cfg["L"]
    circ = create_ising_circuit(L, cfg["J"], cfg["g"], cfg["dt"], cfg["timesteps"], sample_every_step=True)
    sp = DigitalSimParams(observables=[Observable(Z(), s) for s in range(L)], num_traj=cfg["ntraj"], max_bond_dim=cfg["chi"],
                          svd_threshold=cfg["thr"], sample_layers=True, random_seed=cfg["seed"])
    noise = NoiseModel([{"name": n, "sites": [i], "strength": g} for i in range(L) for (n, g) in cfg["noise"]])
    res = Simulator(show_progress=False, kernels=NumpyKernels()).run(State(L, initial="zeros"), circ, sp, noise)
    g = np.load(os.path.join(golden_dir, "dig_ising6_nn_noisy.npz"))
    got = np.stack([res.trajectories[k] for k in range(L)], axis=1)               # [traj, obs, cols]
    assert np.max(np.abs(got - g["results"])) < 1e-10
    assert np.allclose(res.max_bond, g["diagnostics"][:, 1, :].mean(axis=0))
    # noise-free: one trajectory, equals the dense evolution
    sp2 = DigitalSimParams(observables=[Observable(Z(), s) for s in range(L)] + [Observable(X(), 2)], max_bond_dim=None, svd_threshold=1e-12)
    res2 = Simulator(kernels=NumpyKernels()).run(State(L, initial="zeros"), create_ising_circuit(L, 1.0, 0.7, 0.2, 3), sp2)
    x, z = np.array([[0, 1], [1, 0]], dtype=complex), np.diag([1.0 + 0j, -1.0])

    def op_on(m, sites):
        out = np.eye(1)
        mats = {s: m_ for s, m_ in zip(sites, m)}
        for s in range(L):
            out = np.kron(out, mats.get(s, np.eye(2)))                               # site 0 = most significant (MPS order)
        return out
    psi = np.zeros(2 ** L, dtype=complex)
    psi[0] = 1
    a, b = -2 * 0.2 * 0.7, -2 * 0.2 * 1.0
    rx = np.cos(a / 2) * np.eye(2) - 1j * np.sin(a / 2) * x
    for _ in range(3):
        for s in range(L):
            psi = op_on([rx], [s]) @ psi
        for s in range(L - 1):
            zz = op_on([z, z], [s, s + 1])
            psi = (np.cos(b / 2) * np.eye(2 ** L) - 1j * np.sin(b / 2) * zz) @ psi
    for s in range(L):
        assert abs(res2.expectation_values[s][-1] - np.vdot(psi, op_on([z], [s]) @ psi).real) < 1e-9
    assert abs(res2.expectation_values[L][-1] - np.vdot(psi, op_on([x], [2]) @ psi).real) < 1e-9
    assert len(res2.trajectories[0]) == 1


def test_simulator_dynamic_tdvp_mode(golden_dir):
    """AnalogSimParams(tdvp_mode="dynamic") through Simulator.run: the uncapped fixture of the unmodified reference is reproduced; with a
    cap the run warns that the reference's hand-over at the cap is not reproduced."""
    import warnings

    from oracle.golden_cases import DYNAMIC_CASES
    from yaqs_b200 import ZZ
    cfg = DYNAMIC_CASES["heis6_dynamic_nocap_order2"]
    L = cfg["L"]
    obs = [Observable(Z(), s) for s in range(L)] + [Observable(ZZ(), [s, s + 1]) for s in range(L - 1)]
    noise = NoiseModel([{"name": n, "sites": [i], "strength": g} for i in range(L) for (n, g) in cfg["noise"]])
    sp = AnalogSimParams(observables=obs, elapsed_time=cfg["dt"] * cfg["steps"], dt=cfg["dt"], num_traj=cfg["ntraj"], max_bond_dim=None,
                         svd_threshold=cfg["thr"], krylov_tol=1e-4, order=2, sample_timesteps=True, random_seed=cfg["seed"], tdvp_mode="dynamic")
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        res = Simulator(show_progress=False, kernels=NumpyKernels()).run(State(L, initial="Neel"), Hamiltonian.heisenberg(L, 1.0, 1.0, 1.0), sp, noise)
    g = np.load(os.path.join(golden_dir, "heis6_dynamic_nocap_order2.npz"))
    order = sp.observable_sorted_indices
    for k in range(len(obs)):
        assert np.max(np.abs(np.asarray(res.trajectories[k]) - g["results"][:, order[k], :])) < 1e-10
    sp2 = AnalogSimParams(observables=obs[:L], elapsed_time=0.2, dt=0.1, num_traj=2, max_bond_dim=4, order=1, random_seed=1, tdvp_mode="dynamic")
    with pytest.warns(UserWarning, match="projector-splitting"):
        res2 = Simulator(show_progress=False, kernels=NumpyKernels()).run(State(L, initial="x+"), Hamiltonian.ising(L, 1.0, 0.8), sp2, noise)
    assert np.all(res2.max_bond <= 4 * 4) and np.all(np.isfinite(res2.expectation_values[0]))


def test_digital_get_state_only_run():
    """DigitalSimParams(get_state=True) without observables or shots (simulator.py:1282-1290,1714-1716): the noise-free final state comes
    back as Result.output_state and carries the observables of an observable run; a noisy model is rejected."""
    from yaqs_b200 import DigitalSimParams
    from yaqs_b200.circuits import create_ising_circuit
    L = 5
    circ = create_ising_circuit(L, 1.0, 0.6, 0.2, 3)
    sim = Simulator(show_progress=False, kernels=NumpyKernels())
    res = sim.run(State(L, initial="zeros"), circ, DigitalSimParams(get_state=True, max_bond_dim=None, svd_threshold=1e-12))
    vec = res.output_state.mps.to_vector()
    assert abs(np.vdot(vec, vec) - 1.0) < 1e-12
    obs = sim.run(State(L, initial="zeros"), circ, DigitalSimParams(observables=[Observable(Z(), s) for s in range(L)], max_bond_dim=None,
                                                                    svd_threshold=1e-12))
    psi = vec.reshape((2,) * L)
    for s in range(L):
        z = sum((1 - 2 * i) * float(np.sum(np.abs(np.take(psi, i, axis=s)) ** 2)) for i in (0, 1))
        assert abs(z - obs.expectation_values[s][-1]) < 1e-10
    with pytest.raises(ValueError, match="No output specified"):
        sim.run(State(L), circ, DigitalSimParams())
    with pytest.raises(ValueError, match="noisy circuit simulation"):
        sim.run(State(L), circ, DigitalSimParams(get_state=True), NoiseModel([{"name": "pauli_x", "sites": [0], "strength": 0.1}]))


def test_open_advance_result_equals_run():
    """Simulator.open / AnalogRun.advance / result (step-wise, per-step host I/O) == Simulator.run, bit for bit; the per-step I/O is
    counted (2 trajectory + 2 sample uniforms per trajectory and step in, one result + diagnostics column out)."""
    length = 5
    sp = AnalogSimParams(observables=[Observable(Z(), s) for s in range(length)], elapsed_time=0.5, dt=0.1, num_traj=6, max_bond_dim=4,
                         svd_threshold=1e-6, order=2, sample_timesteps=True, random_seed=7)
    noise = NoiseModel([{"name": n, "sites": [i], "strength": 0.3} for i in range(length) for n in ["lowering", "pauli_z"]])
    args = (State(length, initial="x+"), Hamiltonian.ising(length, J=1, g=0.5), sp, noise)
    ref = Simulator(show_progress=False, kernels=NumpyKernels()).run(*args)
    run = Simulator(show_progress=False, kernels=NumpyKernels(), lanes=2, batch_size=6).open(*args)
    assert run.n_times == 6
    assert run.advance(1) == 0 and run.advance(2) == 2            # index 0 = initial measurement
    partial = run.result()
    assert np.array_equal(partial.trajectories[0][:, :3], ref.trajectories[0][:, :3]) and not partial.trajectories[0][:, 3:].any()
    assert run.advance(None) == 5 and run.advance(1) == 5
    res = run.result()
    for x, y in zip(res.trajectories, ref.trajectories):
        assert np.array_equal(x, y)
    assert np.array_equal(res.max_bond, ref.max_bond)
    h2d, d2h = run.io_bytes()
    assert h2d == 6 * 6 * 4 * 8 and d2h == 6 * 6 * (length + 3) * 8
    with pytest.raises(ValueError, match="resident"):
        Simulator(show_progress=False, kernels=NumpyKernels(), batch_size=2).open(*args)


@pytest.mark.parametrize("bit,etype,text", [(1, np.linalg.LinAlgError, "SVD did not converge"), (2, ValueError, "infs or NaNs"),
                                            (4, ValueError, "Jump probability weights are zero or non-finite"), (8, AssertionError, "Measurement should be real")])
def test_status_bits_become_reference_exceptions(bit, etype, text):
    """SURVEY 8b error row: every device status bit maps to the exception type / message of the reference (stochastic_process.py:
    179-187, mps.py:1233, the scipy.linalg.svd failure modes) and the complete Result travels with it; on_trajectory_error="flag"
    returns the Result instead, flagged trajectories excluded from the means."""
    from yaqs_b200 import engine as E
    from yaqs_b200.simulator import TrajectoryError
    length = 4
    sp = AnalogSimParams(observables=[Observable(Z(), s) for s in range(length)], elapsed_time=0.2, dt=0.1, num_traj=5, max_bond_dim=4, order=1,
                         random_seed=3)
    noise = NoiseModel([{"name": "pauli_z", "sites": [i], "strength": 0.1} for i in range(length)])
    args = (State(length, initial="x+"), Hamiltonian.ising(length, J=1, g=0.5), sp, noise)
    orig = E.BatchedTJM.run_iter

    def flagged(self, *a, **k):
        for j in orig(self, *a, **k):
            self.status[1] |= bit                               # what a kernel does when trajectory 1 of the batch goes wrong
            yield j

    E.BatchedTJM.run_iter = flagged
    try:
        with pytest.raises(etype, match=text) as info:
            Simulator(show_progress=False, kernels=NumpyKernels()).run(*args)
        exc = info.value
        assert isinstance(exc, TrajectoryError) and exc.trajectories == (1,) and int(exc.status[1]) == bit
        assert exc.result is not None and exc.result.trajectories[0].shape == (5, 3)
        res = Simulator(show_progress=False, kernels=NumpyKernels(), on_trajectory_error="flag").run(*args)
    finally:
        E.BatchedTJM.run_iter = orig
    clean = Simulator(show_progress=False, kernels=NumpyKernels()).run(*args)
    assert list(res.status) == [0, bit, 0, 0, 0]
    keep = [0, 2, 3, 4]
    for k in range(length):
        assert np.allclose(res.expectation_values[k], clean.trajectories[k][keep].mean(axis=0), atol=1e-15)


def test_digital_tdvp_mode_only_matters_for_window_gates():
    """A nearest-neighbour circuit never takes a TDVP window, so tdvp_mode="1site" runs (the reference raises only inside
    apply_two_qubit_gate_tdvp, digital_tjm.py:324-340); a periodic circuit in gate_mode="tdvp" does take one and raises there."""
    from yaqs_b200 import DigitalSimParams
    from yaqs_b200.circuits import create_ising_circuit
    L = 4
    obs = [Observable(Z(), s) for s in range(L)]
    sim = Simulator(show_progress=False, kernels=NumpyKernels())
    a = sim.run(State(L, initial="zeros"), create_ising_circuit(L, 1.0, 0.6, 0.2, 2), DigitalSimParams(observables=obs, tdvp_mode="1site"))
    b = sim.run(State(L, initial="zeros"), create_ising_circuit(L, 1.0, 0.6, 0.2, 2), DigitalSimParams(observables=obs))
    for x, y in zip(a.expectation_values, b.expectation_values):
        assert np.array_equal(x, y)
    with pytest.raises(ValueError, match='only supports tdvp_mode="2site"'):
        sim.run(State(L, initial="zeros"), create_ising_circuit(L, 1.0, 0.6, 0.2, 1, periodic=True),
                DigitalSimParams(observables=obs, tdvp_mode="1site", gate_mode="tdvp"))


def test_user_tensors_keep_their_norm():
    """The reference keeps user tensors as given (core/data_structures/mps.py:112-114).  Here they are re-gauged (right-canonical,
    centre 0) WITHOUT rescaling: the represented vector is unchanged, an unnormalised input stays unnormalised, and MPS.normalize()
    is what divides the norm out."""
    from yaqs_b200.containers import MPS
    rng = np.random.default_rng(3)
    bonds = [1, 2, 4, 2, 1]
    tensors = [rng.standard_normal((2, bonds[i], bonds[i + 1])) + 1j * rng.standard_normal((2, bonds[i], bonds[i + 1])) for i in range(4)]

    def dense(ts):
        v = np.ones((1, 1), dtype=np.complex128)
        for t in ts:
            v = np.einsum("xl,slr->xsr", v, t).reshape(-1, t.shape[2])
        return v.reshape(-1)

    ref = dense(tensors)
    mps = MPS(4, tensors=[t.copy() for t in tensors])
    assert mps.orthogonality_center == 0
    assert np.max(np.abs(dense(mps.tensors) - ref)) < 1e-12 * np.abs(ref).max()
    assert abs(np.linalg.norm(mps.tensors[0]) - np.linalg.norm(ref)) < 1e-12 * np.linalg.norm(ref)
    for t in mps.tensors[1:]:                                  # right isometries
        m = t.transpose(1, 0, 2).reshape(t.shape[1], -1)
        assert np.max(np.abs(m @ m.conj().T - np.eye(t.shape[1]))) < 1e-12
    mps.normalize()
    assert abs(np.linalg.norm(dense(mps.tensors)) - 1.0) < 1e-12


def test_lanes_first_mode_runs_first_item_then_calls_between():
    """Lanes.map(serial="first") (bench.py's instrumented step): the first item alone, the callback, then the rest; on a CPU device
    (no streams) it degrades to the serial order with the callback at the end."""
    from yaqs_b200.lanes import Lanes
    order = []
    lanes = Lanes("cpu", 3)
    out = lanes.map(lambda x: (order.append(x), x * 2)[1], [1, 2, 3], serial="first", between=lambda: order.append("between"))
    assert out == [2, 4, 6]
    assert order[0] == 1 and "between" in order and order.index("between") >= 1


def test_run_rejects_what_the_reference_rejects_with_its_messages():
    """Messages recorded from the unmodified reference's Simulator.run / validate_noise_model_for_run (simulator.py:817,1283-1308,
    noise_model.py:695-780) on the same calls."""
    import yaqs_b200 as Y
    sim = Y.Simulator(show_progress=False, kernels=NumpyKernels())
    obs = [Y.Observable("z", 0)]
    ap = Y.AnalogSimParams(observables=obs, elapsed_time=0.1, dt=0.1, num_traj=2)
    with pytest.raises(ValueError, match=r"State.length=4 does not match Hamiltonian.length=3\."):
        sim.run(Y.State(4), Y.Hamiltonian.ising(3, 1, 0.5), ap)
    with pytest.raises(ValueError, match=r"Process site index 7 is out of range for length 3\."):
        sim.run(Y.State(3), Y.Hamiltonian.ising(3, 1, 0.5), ap, Y.NoiseModel([{"name": "lowering", "sites": [7], "strength": 0.1}]))
    with pytest.raises(TypeError, match="Circuit simulation requires a QuantumCircuit operator."):
        sim.run(Y.State(3), Y.Hamiltonian.ising(3, 1, 0.5), Y.DigitalSimParams(observables=obs))
    with pytest.raises(ValueError, match="No output specified: set observables, shots, and/or get_state."):
        sim.run(Y.State(3), Y.create_ising_circuit(3, 1, 0.5, 0.1, 1), Y.DigitalSimParams())
    with pytest.raises(AssertionError, match="Observable acting on non-existing site: 5"):
        sim.run(Y.State(3), Y.Hamiltonian.ising(3, 1, 0.5), Y.AnalogSimParams(observables=[Y.Observable("z", 5)], elapsed_time=0.1, dt=0.1))
    long_range = Y.NoiseModel([{"name": "crosstalk_xx", "sites": [0, 2], "strength": 0.1}])
    with pytest.raises(ValueError, match=r"Digital TJM does not support non-adjacent / factorized two-site noise \(process 'crosstalk_xx' on sites \[0, 2\]\)"):
        sim.run(Y.State(3), Y.create_ising_circuit(3, 1, 0.5, 0.1, 1), Y.DigitalSimParams(observables=obs, num_traj=2), long_range)
    sched = Y.NoiseModel([{"name": "lowering", "sites": [0], "strength": 0.1}], scheduled_jumps=[{"time": 0.05, "sites": [1], "name": "pauli_x"}])
    with pytest.raises(ValueError, match=r"Scheduled jump time 0.05 is not on the simulation time grid"):
        sim.run(Y.State(3), Y.Hamiltonian.ising(3, 1, 0.5), ap, sched)
