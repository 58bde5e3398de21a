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
from .rng import numpy_tables, philox_tables


def validate_digital_noise(noise, length: int) -> None:
    """noise_model.py:668-757, digital branch: two-site processes must act on adjacent, ascending sites (:724-730)."""
    for pr in noise.processes:
        for s in pr["sites"]:
            if s >= length:
                raise ValueError(f"Noise process '{pr['name']}' acts on site {s} but the chain has {length} sites.")
        if len(pr["sites"]) == 2 and pr["sites"][1] - pr["sites"][0] != 1:
            raise ValueError("Digital simulation supports two-site noise on adjacent sites only.")


def run_digital(simulator, initial_state, operator, sp, noise, result):
    from .simulator import shard_range
    if sp.shots is not None:
        raise NotImplementedError("measure_shots (DigitalSimParams(shots=...)) is a SURVEY.md 8(f) 'next' row of the B200 path")
    if not sp.observables:
        raise ValueError("No output specified: observables must be set (get_state-only runs are outside the B200 hot path).")
    compiled = operator if isinstance(operator, CompiledCircuit) else compile_circuit(operator)
    length = initial_state.length
    if noise is not None:
        validate_digital_noise(noise, length)
    noiseless = noise is None or all(pr["strength"] == 0 for pr in noise.processes)
    ntraj = 1 if noiseless else sp.num_traj                                                   # simulator.py:1735-1745
    sorted_obs = sp.sorted_observables
    to_sorted = sp.observable_sorted_indices
    obs = engine_observables(sorted_obs, length)
    ep = EngineParams(dt=1.0, n_steps=1, order=1, max_bond_dim=sp.max_bond_dim, svd_threshold=sp.svd_threshold, trunc_mode=sp.trunc_mode,
                      krylov_tol=sp.krylov_tol, tdvp_mode=sp.tdvp_mode, tdvp_sweeps=sp.tdvp_sweeps, sample_timesteps=True, digital=True)
    if ep.tdvp_mode != "2site":
        raise ValueError(f'apply_two_qubit_gate_tdvp only supports tdvp_mode="2site"; got {ep.tdvp_mode!r}.')
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
    wave = simulator._wave_size(len(mine), length, ep, mpo, init, dev)
    K.launch_count(reset=True)
    for w0 in range(0, len(mine), max(wave, 1)):
        idx = list(mine[w0: w0 + wave])
        eng = BatchedTJM(K, length, len(idx), mpo, ep, procs, obs, init)
        if simulator.rng_mode == "numpy":
            tu, _ = numpy_tables(sp.random_seed, idx, max(n2q, 1), 1)
            tu = torch.from_numpy(tu).to(dev)
        else:
            tu, _ = philox_tables(K, sp.random_seed, idx, max(n2q, 1), 1, dev)
        log = [] if simulator.collect_log else None
        r, dg = eng.run_digital(compiled.layers, tu, sp.sample_layers, sp.gate_mode, log)
        res_local[w0: w0 + len(idx)] = r.cpu().numpy()
        diag_local[w0: w0 + len(idx)] = dg.cpu().numpy()
        status_local[w0: w0 + len(idx)] = eng.status.cpu().numpy()
        for k, v in eng.stats.items():
            stats[k] = stats.get(k, 0) + v
        if log_all is not None:
            log_all.append((idx, log))
        del eng
    stats["kernel_launches"] = K.launch_count()
    simulator.last_log = log_all
    times = np.arange(ncol, dtype=np.float64)
    simulator._finish(result, sp, res_local, diag_local, status_local, mine, ntraj, len(obs), ncol, to_sorted, times, stats, dev)
    return result
