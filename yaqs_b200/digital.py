"""``Simulator.run`` for circuits: the digital TJM (digital/digital_tjm.py:636-749) as GPU batches.

Reference flow: ``Simulator._run_digital_sim`` (simulator.py:1684-1835) compiles the circuit once and fans
``digital_tjm`` out over the process pool; here every wave of trajectories advances through the compiled layers in
lock-step on the device (``engine.BatchedTJM.run_digital``).  The circuit may be a ``yaqs_b200.circuits.Circuit``
(qiskit-free), an already compiled ``CompiledCircuit``, or a qiskit ``QuantumCircuit`` when qiskit is importable.
"""
from __future__ import annotations

import numpy as np
import torch

from .circuits import CompiledCircuit, compile_circuit
from .engine import BatchedTJM, EngineParams
from .params import engine_observables
from .rng import STREAM_SHOTS, numpy_tables, philox_source


def validate_digital_noise(noise, length: int) -> None:
    """noise_model.py:668-757, digital branch (processes, then the scheduled-jump refusal), the reference's messages."""
    from .simulator import validate_noise_model_for_run
    validate_noise_model_for_run(noise, length, None, is_digital=True)


def run_digital(simulator, initial_state, operator, sp, noise, result):
    from .simulator import shard_range
    wants_obs, wants_shots = bool(sp.observables), sp.shots is not None
    if not wants_obs and not wants_shots and not sp.get_state:
        raise ValueError("No output specified: set observables, shots, and/or get_state.")             # simulator.py:1282-1290
    compiled = operator if isinstance(operator, CompiledCircuit) else compile_circuit(operator)
    length = initial_state.length
    if noise is not None:
        validate_digital_noise(noise, length)
    noiseless = noise is None or all(pr["strength"] == 0 for pr in noise.processes)
    if not noiseless and sp.get_state:
        raise ValueError("Cannot return state in noisy circuit simulation due to stochastics.")          # simulator.py:1714-1716
    ntraj, shots_of = plan_digital_shots(sp, noisy=not noiseless)
    sorted_obs = sp.sorted_observables
    to_sorted = sp.observable_sorted_indices
    obs = engine_observables(sorted_obs, length)
    ep = EngineParams(dt=1.0, n_steps=1, order=1, max_bond_dim=sp.max_bond_dim, svd_threshold=sp.svd_threshold, trunc_mode=sp.trunc_mode,
                      krylov_tol=sp.krylov_tol, tdvp_mode=sp.tdvp_mode, tdvp_sweeps=sp.tdvp_sweeps, sample_timesteps=True, digital=True)
    mpo = [np.eye(2, dtype=np.complex128).reshape(2, 2, 1, 1) for _ in range(length)]      # bond-1 layout for generator windows
    procs = list(noise.processes) if noise is not None else None
    init = initial_state.mps.tensors
    n2q = sum(len(l.even_two_qubit_gates) + len(l.odd_two_qubit_gates) for l in compiled.layers)
    mid = compiled.num_mid_measurements if sp.sample_layers else 0
    ncol = mid + 2 if sp.sample_layers else 1

    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    mine = shard_range(ntraj, rank, world)
    K = simulator._get_kernels()
    dev = K.device
    res_local = np.zeros((len(mine), len(obs), ncol))
    diag_local = np.zeros((len(mine), 3, ncol))
    status_local = np.zeros(len(mine), dtype=np.int32)
    stats: dict = {}
    log_all = [] if simulator.collect_log else None
    measurements_local: dict[int, dict[int, int]] = {}
    bond_factor = BatchedTJM.mpo_route_bond_factor(compiled.layers, sp.gate_mode, length)      # head-room of the gate-MPO route
    wave = simulator._wave_size(len(mine), length, ep, mpo, init, dev, bond_factor=bond_factor)
    K.launch_count(reset=True)
    for w0 in range(0, len(mine), max(wave, 1)):
        idx = list(mine[w0: w0 + wave])
        eng = BatchedTJM(K, length, len(idx), mpo, ep, procs, obs, init, bond_factor=bond_factor)
        eng.on_sample = getattr(simulator, "on_sample", None)
        if simulator.rng_mode == "numpy":
            tu, _ = numpy_tables(sp.random_seed, idx, max(n2q, 1), 1)
            tu = torch.from_numpy(tu).to(dev)
        else:
            tu = None
            eng.philox = philox_source(sp.random_seed, idx, dev)
        log = [] if simulator.collect_log else None
        if torch.device(dev).type == "cuda":               # a side stream: the legacy default stream cannot hold CUDA-graph captures (engine._run_graphed)
            side, cur = torch.cuda.Stream(dev), torch.cuda.current_stream(dev)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                r, dg = eng.run_digital(compiled.layers, tu, sp.sample_layers, sp.gate_mode, log)
            cur.wait_stream(side)
        else:
            r, dg = eng.run_digital(compiled.layers, tu, sp.sample_layers, sp.gate_mode, log)
        res_local[w0: w0 + len(idx)] = r.cpu().numpy()
        diag_local[w0: w0 + len(idx)] = dg.cpu().numpy()
        status_local[w0: w0 + len(idx)] = eng.status.cpu().numpy()
        if wants_shots:
            per = np.array([shots_of(t) for t in idx], dtype=np.int64)
            if per.max() > 0:                                                                    # never sample 0 shots (digital_tjm.py:746)
                su = shot_tables(K, simulator.rng_mode, sp.random_seed, idx, int(per.max()), length, dev)
                outs = eng.measure_shots(eng.phi, su).cpu().numpy()
                for row, t in enumerate(idx):
                    if per[row] > 0:
                        vals, cnt = np.unique(outs[row, : per[row]], return_counts=True)
                        measurements_local[t] = {int(v): int(c) for v, c in zip(vals, cnt)}
        for k, v in eng.stats.items():
            stats[k] = stats.get(k, 0) + v
        if log_all is not None:
            log_all.append((idx, log))
        if sp.get_state and noiseless:
            from .containers import State
            result.output_state = State(tensors=eng.phi.download(0))
        del eng
    stats["kernel_launches"] = K.launch_count()
    simulator.last_log = log_all
    times = np.arange(ncol, dtype=np.float64)
    simulator._finish(result, sp, res_local, diag_local, status_local, mine, ntraj, len(obs), ncol, to_sorted, times, stats, dev)
    if wants_shots:
        if world > 1:
            gathered: list = [None] * world
            dist.all_gather_object(gathered, measurements_local)
            measurements_local = {k: v for part in gathered for k, v in part.items()}
        # simulator.py:1752-1756,1790-1795: one entry per trajectory when noisy, a single entry otherwise
        result.measurements = [measurements_local.get(t) for t in range(ntraj if not noiseless else 1)]
        aggregate_counts(result)
    return result


def plan_digital_shots(sp, *, noisy: bool):
    """_plan_digital_shots + _per_call_shots (simulator.py:1001-1050, digital_tjm.py:752-768): (number of trajectories,
    trajectory index -> shots to sample from its final state)."""
    wants_obs, wants_shots = bool(sp.observables), sp.shots is not None
    if wants_shots and not wants_obs:
        if noisy:
            return int(sp.shots), (lambda t: 1)                    # one stochastic state per shot (num_traj is ignored)
        return 1, (lambda t: int(sp.shots))                        # the whole budget from the one noiseless state
    ntraj = sp.num_traj if noisy else 1
    if not wants_shots:
        return ntraj, (lambda t: 0)
    if noisy:
        base, rem = divmod(int(sp.shots), int(ntraj))              # budget spread over the observable ensemble
        return ntraj, (lambda t: base + (1 if t < rem else 0))
    return ntraj, (lambda t: int(sp.shots))


def shot_tables(kernels, rng_mode: str, seed, traj_indices, nshots: int, length: int, device) -> torch.Tensor:
    """[B, nshots, L] uniforms for the projective sampling.  The reference draws them from an unseeded Generator
    (mps.py:1312-1313); here they come from the per-trajectory "SHOT" stream, so shots are reproducible under random_seed."""
    nb, n = len(traj_indices), nshots * length
    if rng_mode == "numpy":
        tab = np.empty((nb, n))
        for row, t in enumerate(traj_indices):
            g = np.random.default_rng() if seed is None else np.random.default_rng(np.random.SeedSequence([seed, int(t), STREAM_SHOTS]))
            tab[row] = g.random(n)
        return torch.from_numpy(tab).to(device).view(nb, nshots, length)
    base = int(seed) if seed is not None else int(np.random.SeedSequence().entropy % (2 ** 64))
    nblocks = (n + 1) // 2
    flat = torch.empty((nb, 2 * nblocks), dtype=torch.float64, device=device)
    kernels.philox_uniform(base, torch.as_tensor(np.asarray(traj_indices, dtype=np.int64), device=device), STREAM_SHOTS, nblocks, 0, flat)
    return flat[:, :n].contiguous().view(nb, nshots, length)


def aggregate_counts(result) -> None:
    """result.py:141-152."""
    counts: dict[int, int] = {}
    for m in filter(None, result.measurements):
        for key, value in m.items():
            counts[key] = counts.get(key, 0) + value
    result.counts = dict(sorted(counts.items()))
