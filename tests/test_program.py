"""SimulationProgram against fixtures of the unmodified reference's _execute_program_trajectory
(oracle/make_program_golden.py): per segment the same observables (<= 1e-10) and identical bond diagnostics for every trajectory,
and the same number of doubles drawn from each trajectory's shared stream (merge / hand-off / digital bookkeeping)."""
import os

import numpy as np
import pytest

import yaqs_b200 as Y
from oracle.golden_cases import PROGRAM_CASES
from yaqs_b200.params import prepare_observable_ordering
from yaqs_b200.program import compile_program, plan_units


def build_program(cfg):
    L = cfg["L"]
    obs = []
    for tok in cfg["obs"].split("+"):
        if tok == "z":
            obs += [Y.Observable(Y.Z(), s) for s in range(L)]
        elif tok == "x":
            obs += [Y.Observable(Y.X(), s) for s in range(L)]
        elif tok == "zz":
            obs += [Y.Observable(Y.ZZ(), [s, s + 1]) for s in range(L - 1)]
    procs = [{"name": n, "sites": [i], "strength": g} for i in range(L) for (n, g) in cfg["noise"]]
    noise = Y.NoiseModel(procs) if procs else None
    hams, segs = {}, []
    for seg in cfg["segments"]:
        if seg["kind"] == "A":
            if seg["ham"] not in hams:
                hams[seg["ham"]] = Y.Hamiltonian.ising(L, seg["J"], seg["g"])
            segs.append((hams[seg["ham"]], Y.AnalogSimParams(elapsed_time=seg["T"], dt=seg["dt"], num_traj=cfg["ntraj"], max_bond_dim=seg["chi"],
                                                            svd_threshold=seg["thr"], order=seg["order"], sample_timesteps=seg["sample"])))
        else:
            circ = Y.create_ising_circuit(L, seg["J"], seg["g"], seg["dt"], seg["timesteps"], sample_every_step=seg["sample"])
            segs.append((circ, Y.DigitalSimParams(num_traj=cfg["ntraj"], max_bond_dim=seg["chi"], svd_threshold=seg["thr"], sample_layers=seg["sample"])))
    return Y.SimulationProgram(segs, observables=obs, random_seed=cfg["seed"]), Y.State(L, initial=cfg["init"]), noise


def check_program(name, golden_dir, kernels):
    cfg = PROGRAM_CASES[name]
    g = np.load(os.path.join(golden_dir, f"{name}.npz"))
    prog, state, noise = build_program(cfg)
    res = Y.Simulator(show_progress=False, kernels=kernels).run(state, prog, None, noise)
    assert len(res.segment_results) == int(g["nseg"])
    ntraj = g["res_0"].shape[0]
    pieces = [[] for _ in prog.observables]
    to_sorted = prepare_observable_ordering(list(prog.observables))[1]
    for k, seg in enumerate(res.segment_results):
        want, dg = g[f"res_{k}"], g[f"diag_{k}"]                       # [ntraj, n_obs, cols] (the observables are given in sorted order)
        got = np.zeros_like(want)                                      # fixture rows follow the reference's site-sorted order
        for i in range(len(prog.observables)):
            got[:, to_sorted[i]] = np.asarray(seg.trajectories[i])
        assert got.shape == want.shape, (name, k, got.shape, want.shape)
        assert np.max(np.abs(got - want)) < 1e-10, (name, k, float(np.max(np.abs(got - want))))
        assert np.array_equal(seg.runtime_cost, dg[:, 0].mean(axis=0)) and np.array_equal(seg.max_bond, dg[:, 1].mean(axis=0))
        assert np.array_equal(seg.total_bond, dg[:, 2].mean(axis=0))
        assert seg.segment_index == k and seg.time_offset == pytest.approx(float(g["time_offsets"][k]))
        for i in range(len(prog.observables)):
            pieces[i].append(want[:, to_sorted[i]].mean(axis=0))
    assert np.array_equal(res.stats["trajectory_draws"][:ntraj], g["draws"])
    for i in range(len(prog.observables)):                             # stitched outer result (simulation_program.py:799-842)
        assert np.max(np.abs(res.expectation_values[i] - np.concatenate(pieces[i][:]))) < 1e-10
    assert len(res.times) == len(res.expectation_values[0])


@pytest.mark.parametrize("name", sorted(PROGRAM_CASES))
def test_program_matches_reference_host_twin(golden_dir, name):
    from oracle.np_kernels import NumpyKernels
    check_program(name, golden_dir, NumpyKernels())


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(PROGRAM_CASES))
def test_program_matches_reference_gpu(golden_dir, name):
    from yaqs_b200.kernels import CudaKernels
    check_program(name, golden_dir, CudaKernels("cuda:0"))


def test_plan_units_follow_the_reference_control_flow():
    """The trajectory-independent plan: merges, order-2 hand-offs and sample-stream offsets (simulator.py:295-336,405-427,588-718)."""
    kinds = {}
    for name, cfg in PROGRAM_CASES.items():
        prog, state, noise = build_program(cfg)
        units = plan_units(compile_program(prog, state, noise).instructions)
        kinds[name] = [(u.kind, len(u.instructions), u.continue_trajectory, u.hand_off, u.sample_offset, u.sample_at) for u in units]
    assert kinds["prog_quench_merged_o2"] == [("analog", 2, False, False, 0, None)]
    assert kinds["prog_handoff_o2"] == [("analog", 1, False, True, 0, None), ("analog", 1, True, False, 2, None), ("analog", 1, False, False, 0, None)]
    assert kinds["prog_mixed_digital"] == [("analog", 2, False, False, 0, (2, 4)), ("digital", 1, False, False, 0, None),
                                           ("analog", 1, False, False, 0, None)]


def test_program_validation_errors():
    L = 4
    ham = Y.Hamiltonian.ising(L, 1.0, 0.5)
    sp = Y.AnalogSimParams(elapsed_time=0.1, dt=0.1)
    with pytest.raises(ValueError, match="at least one segment"):
        Y.SimulationProgram([])
    with pytest.raises(TypeError, match="expected AnalogSimParams"):
        Y.SimulationProgram([(ham, Y.DigitalSimParams())])
    with pytest.raises(ValueError, match="observables must be empty"):
        Y.SimulationProgram([(ham, Y.AnalogSimParams(observables=[Y.Observable(Y.Z(), 0)], elapsed_time=0.1, dt=0.1))])
    with pytest.raises(ValueError, match="random_seed must be None"):
        Y.SimulationProgram([(ham, Y.AnalogSimParams(elapsed_time=0.1, dt=0.1, random_seed=1))])
    with pytest.raises(ValueError, match="disagree on sim_params.num_traj"):
        compile_program(Y.SimulationProgram([(ham, Y.AnalogSimParams(elapsed_time=0.1, dt=0.1, num_traj=2)),
                                             (ham, Y.AnalogSimParams(elapsed_time=0.1, dt=0.1, num_traj=3))]), Y.State(L), None)
    sim = Y.Simulator(show_progress=False, kernels=object())
    with pytest.raises(TypeError, match="sim_params must be None"):
        sim.run(Y.State(L), Y.SimulationProgram([(ham, sp)], observables=[Y.Observable(Y.Z(), 0)]), sp)
    with pytest.raises(ValueError, match="No output specified"):
        sim.run(Y.State(L), [(ham, sp)])
    with pytest.raises(TypeError, match="only valid for a SimulationProgram"):
        sim.run(Y.State(L), ham, sp, None, num_traj=3)
    with pytest.raises(ValueError, match="does not match State.length"):
        sim.run(Y.State(L + 1), [(ham, sp)], observables=[Y.Observable(Y.Z(), 0)])
    noisy = Y.NoiseModel([{"name": "lowering", "sites": [0], "strength": 0.1}])
    with pytest.raises(ValueError, match="Cannot return state from a noisy SimulationProgram"):
        sim.run(Y.State(L), [(ham, sp)], None, noisy, get_state=True)


def test_program_get_state_and_checkpoints_noiseless():
    """Noise-free program with get_state on the program and on the first segment: one trajectory, the outer output_state carries the
    last column's observables, the segment checkpoint those of its own last column; equal Hamiltonians as separate segments give
    the same physics as one long standalone run (order 1: exactly the same sequence of sweeps)."""
    from oracle.np_kernels import NumpyKernels
    L = 5
    obs = [Y.Observable(Y.Z(), s) for s in range(L)]
    ham = Y.Hamiltonian.ising(L, 1.0, 0.7)

    def params(t, **kw):
        return Y.AnalogSimParams(elapsed_time=t, dt=0.1, order=1, max_bond_dim=8, svd_threshold=1e-12, **kw)
    sim = Y.Simulator(show_progress=False, kernels=NumpyKernels())
    res = sim.run(Y.State(L, initial="x+"), [(ham, params(0.2, get_state=False)), (ham, params(0.3))], observables=obs, get_state=True)
    assert [len(s.trajectories[0]) for s in res.segment_results] == [1, 1]
    whole = sim.run(Y.State(L, initial="x+"), ham, Y.AnalogSimParams(observables=obs, elapsed_time=0.5, dt=0.1, order=1, max_bond_dim=8,
                                                                     svd_threshold=1e-12, get_state=True))
    for k in range(L):
        stitched = res.expectation_values[k]
        assert len(stitched) == 3 + 4 and stitched[2] == stitched[3]           # the junction appears in both segments
        assert np.max(np.abs(np.concatenate([stitched[:3], stitched[4:]]) - whole.expectation_values[k])) < 1e-12
    vec = res.output_state.mps.to_vector()
    want = whole.output_state.mps.to_vector()
    assert abs(abs(np.vdot(vec, want)) - 1.0) < 1e-10
    assert np.allclose(res.times, [0.0, 0.1, 0.2, 0.2, 0.3, 0.4, 0.5])
    # a checkpoint request on a segment breaks the merge and stores that segment's physical end state
    res2 = sim.run(Y.State(L, initial="x+"), [(ham, params(0.2, get_state=False)), (ham, params(0.3))], observables=obs)
    assert res2.segment_results[0].output_state is None and res2.output_state is None


def _shots_program_run():
    from oracle.np_kernels import NumpyKernels
    L = 4
    noise = Y.NoiseModel([{"name": "pauli_x", "sites": [i], "strength": 0.3} for i in range(L)])
    circ = Y.create_ising_circuit(L, 1.0, 0.5, 0.1, 1)
    segs = [(Y.Hamiltonian.ising(L, 1.0, 0.5), Y.AnalogSimParams(elapsed_time=0.2, dt=0.1, order=2, max_bond_dim=4)),
            (circ, Y.DigitalSimParams(max_bond_dim=4, shots=10))]
    return Y.Simulator(show_progress=False, kernels=NumpyKernels()).run(Y.State(L, initial="zeros"), segs, None, noise, random_seed=3, num_traj=7)


def test_program_shots_only_runs_one_trajectory_per_shot():
    """simulator.py:1344-1353: a noisy program with shots and no observables executes max(shots) complete trajectories, one shot each;
    the outer counts are the last shot-recording segment's (simulation_program.py:784-796)."""
    res = _shots_program_run()
    dig = res.segment_results[1]
    assert len(dig.measurements) == 10 and all(sum(m.values()) == 1 for m in dig.measurements)
    assert sum(res.counts.values()) == 10 and res.counts == dig.counts
    assert dig.max_bond is None and res.expectation_values == [] and res.times is None
    assert _shots_program_run().counts == res.counts                         # reproducible under random_seed


def _program_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.np_kernels import NumpyKernels
        prog, state, noise = build_program(PROGRAM_CASES["prog_mixed_digital"])
        res = Y.Simulator(show_progress=False, kernels=NumpyKernels()).run(state, prog, None, noise)
        shots = _shots_program_run()
        if rank == 0:
            out.put(([np.asarray(s.trajectories[0]) for s in res.segment_results], np.asarray(res.expectation_values[0]), shots.counts))
    finally:
        dist.destroy_process_group()


def test_program_two_rank_gloo_sharding_matches_single_process(golden_dir):
    """N > 1 on CPU (gloo, world_size 2): whole program trajectories are sharded by index; per-segment sums take one all-reduce."""
    import multiprocessing as mp
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_program_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    trajs, mean0, counts = q.get(timeout=600)
    for p in procs:
        p.join(timeout=600)
        assert p.exitcode == 0
    from oracle.np_kernels import NumpyKernels
    prog, state, noise = build_program(PROGRAM_CASES["prog_mixed_digital"])
    single = Y.Simulator(show_progress=False, kernels=NumpyKernels()).run(state, prog, None, noise)
    for a, seg in zip(trajs, single.segment_results):
        assert np.array_equal(a, np.asarray(seg.trajectories[0]))
    assert np.max(np.abs(mean0 - single.expectation_values[0])) < 1e-14
    assert counts == _shots_program_run().counts


def test_piecewise_hamiltonian_equals_the_merged_program_fixture(golden_dir):
    """Hamiltonian.piecewise (hamiltonian.py:176-232) expands to one MPO per interval (simulation_program.py:490-547), which is what
    the reference's executor builds when it merges the two segments of prog_quench_merged_o2: a standalone run of the piecewise
    Hamiltonian and a one-segment program of it reproduce that fixture's columns (trajectory stream and sample streams coincide)."""
    from oracle.np_kernels import NumpyKernels
    cfg = PROGRAM_CASES["prog_quench_merged_o2"]
    g = np.load(os.path.join(golden_dir, "prog_quench_merged_o2.npz"))
    L = cfg["L"]
    a, b = cfg["segments"]
    pw = Y.Hamiltonian.piecewise([(Y.Hamiltonian.ising(L, a["J"], a["g"]), a["T"]), (Y.Hamiltonian.ising(L, b["J"], b["g"]), b["T"])])
    assert pw.is_piecewise and pw.duration == pytest.approx(0.5) and len(pw.pieces) == 2
    with pytest.raises(ValueError, match="no single static operator"):
        pw.mpo
    prog, state, noise = build_program(cfg)
    obs = list(prog.observables)
    to_sorted = prepare_observable_ordering(obs)[1]
    want = np.concatenate([g["res_0"], g["res_1"][:, :, 1:]], axis=2)                       # [traj, sorted obs, 6 columns]
    sim = Y.Simulator(show_progress=False, kernels=NumpyKernels())
    sp = Y.AnalogSimParams(observables=obs, elapsed_time=0.5, dt=0.1, num_traj=cfg["ntraj"], max_bond_dim=a["chi"], svd_threshold=a["thr"],
                           order=2, sample_timesteps=True, random_seed=cfg["seed"])
    res = sim.run(state, pw, sp, noise)
    for i in range(len(obs)):
        assert np.max(np.abs(np.asarray(res.trajectories[i]) - want[:, to_sorted[i]])) < 1e-10
    seg_sp = Y.AnalogSimParams(elapsed_time=0.5, dt=0.1, num_traj=cfg["ntraj"], max_bond_dim=a["chi"], svd_threshold=a["thr"], order=2)
    res2 = sim.run(state, [(pw, seg_sp)], None, noise, observables=obs, random_seed=cfg["seed"])
    for i in range(len(obs)):
        assert np.max(np.abs(np.asarray(res2.segment_results[0].trajectories[i]) - want[:, to_sorted[i]])) < 1e-10
    with pytest.raises(ValueError, match="must sum to elapsed_time"):
        sim.run(state, pw, Y.AnalogSimParams(observables=obs, elapsed_time=0.4, dt=0.1, order=2), noise)
    with pytest.raises(ValueError, match="integer multiple of dt"):
        sim.run(state, Y.Hamiltonian.piecewise([(Y.Hamiltonian.ising(L, 1.0, 0.5), 0.25)]),
                Y.AnalogSimParams(observables=obs, elapsed_time=0.3, dt=0.1, order=2), noise)
    with pytest.raises(ValueError, match="nested piecewise"):
        Y.Hamiltonian.piecewise([(pw, 0.1)])


@pytest.mark.parametrize("rng_mode", ["numpy", "philox"])
def test_program_results_do_not_depend_on_wave_size(rng_mode):
    """Trajectories of a program are keyed by their global index in both RNG modes (numpy streams or the Philox kernel), so running them
    as one wave or as waves of two gives the same per-trajectory rows (results independent of batching / sharding)."""
    from oracle.np_kernels import NumpyKernels
    prog, state, noise = build_program(PROGRAM_CASES["prog_mixed_digital"])
    runs = []
    for batch in (None, 2, 3):
        sim = Y.Simulator(show_progress=False, kernels=NumpyKernels(), rng_mode=rng_mode, batch_size=batch)
        runs.append(sim.run(state, prog, None, noise))
    for other in runs[1:]:
        for a, b in zip(runs[0].segment_results, other.segment_results):
            for ta, tb in zip(a.trajectories, b.trajectories):
                assert np.array_equal(np.asarray(ta), np.asarray(tb))
        assert np.array_equal(runs[0].stats["trajectory_draws"], other.stats["trajectory_draws"])
