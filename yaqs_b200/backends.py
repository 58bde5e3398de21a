"""The reference's per-trajectory backend boundary, kept importable.

The reference hands ``backend(args)`` with ``args = (traj_idx, initial MPS, NoiseModel | None, sim_params, operator)`` to a forked
worker (``core/parallel_utils.py:331-390``, called from ``simulator.py:1611-1633`` / ``:1759-1824``).  This module keeps those three
callables -- ``analog_tjm_2`` / ``analog_tjm_1`` (``analog/analog_tjm.py:206,369``) and ``digital_tjm`` (``digital/digital_tjm.py:636``)
-- with the same argument tuple and return shapes, as batches of ONE trajectory, and adds the entry the B200 ``Simulator`` is built
around: ``run_batch``, the same call for a list of trajectory indices advanced in lock step by the CUDA kernels.

There is no CPU fallback: the kernels object is the CUDA library unless a test injects another implementation of the same
entry points with ``set_default_kernels``.
"""
from __future__ import annotations

from typing import Sequence

import numpy as np
import torch

from .containers import MPS, Hamiltonian
from .engine import BatchedTJM, EngineParams
from .noise import NoiseModel
from .params import AnalogSimParams, DigitalSimParams, engine_observables
from .rng import numpy_tables, philox_tables

_DEFAULT_KERNELS = None


def set_default_kernels(kernels) -> None:
    """Kernels object used when a backend is called without ``kernels=`` (None: CudaKernels on the current CUDA device)."""
    global _DEFAULT_KERNELS
    _DEFAULT_KERNELS = kernels


def _kernels(kernels=None):
    if kernels is not None:
        return kernels
    global _DEFAULT_KERNELS
    if _DEFAULT_KERNELS is None:
        from .kernels import CudaKernels          # raises without the built library / a CUDA device
        _DEFAULT_KERNELS = CudaKernels(f"cuda:{torch.cuda.current_device()}" if torch.cuda.is_available() else "cuda:0")
    return _DEFAULT_KERNELS


def _tensors(state) -> list[np.ndarray]:
    mps = getattr(state, "mps", state)            # State wraps an MPS; an MPS is taken as is
    return [np.asarray(t, dtype=np.complex128) for t in mps.tensors]


def _mpo_tensors(operator) -> list[np.ndarray]:
    if isinstance(operator, Hamiltonian):
        operator = operator.mpo
    return operator.tensors if isinstance(operator.tensors, list) else list(operator.tensors)


def _interval_mpos(operator, n_intervals: int):
    """``HamiltonianOperator = MPO | tuple[MPO, ...]`` (analog_tjm.py:40-48): (first MPO's tensors, per-interval tensor lists | None)."""
    if isinstance(operator, (list, tuple)):
        if len(operator) != n_intervals:
            raise ValueError(f"one MPO per analog interval expected ({n_intervals}), got {len(operator)}")
        by_id: dict[int, list] = {}
        per = [by_id.setdefault(id(m), _mpo_tensors(m)) for m in operator]       # same MPO object -> same list (engine caches by id)
        return per[0], (per if len(by_id) > 1 else None)
    return _mpo_tensors(operator), None


def _generator_table(rng: np.random.Generator, n_times: int) -> np.ndarray:
    """An externally managed numpy Generator as a uniform table: a jump test draws ``rng.random()`` and, if it jumps, one more
    double inside ``rng.choice`` (stochastic_process.py:190-245), so 2 * n_times consecutive doubles replay any history.
    ``run_batch`` rewinds the Generator afterwards and advances it by the doubles the trajectory really consumed, so a caller
    that threads one Generator through several calls (SimulationProgram execution, simulator.py:610) sees the reference's state."""
    if not isinstance(rng, np.random.Generator):
        raise TypeError("rng must be a numpy.random.Generator (the batched path replays its doubles from a table)")
    return rng.random(2 * n_times)


def run_batch(traj_indices: Sequence[int], initial_state, noise_model: NoiseModel | None, sim_params: AnalogSimParams, operator, *,
              kernels=None, rng_mode: str = "numpy", trajectory_rngs: Sequence[np.random.Generator] | None = None,
              sample_timestep_offset: int = 0, return_states: bool = False, sample_at: Sequence[int] | None = None,
              continue_trajectory: bool = False, return_trajectory_state: bool = False, use_trajectory_rng_for_final_sample: bool = False):
    """Analog TJM for the trajectories ``traj_indices`` as one lock-step batch.

    Returns ``(results [B, n_sorted_obs, T | 1], diagnostics [B, 3, T | 1], status [B], final)`` as host arrays; rows follow
    ``sim_params.sorted_observables`` exactly like the reference's per-trajectory buffers (``simulator.py:932-942`` maps them back
    to user order).  ``final``: list of per-trajectory tensor lists when ``return_states`` is set, else None -- the physical
    state at the final time, or the mid-Trotter trajectory state phi with ``return_trajectory_state`` (order 2).  ``sample_at``,
    ``continue_trajectory`` and ``use_trajectory_rng_for_final_sample`` are the segment hand-off options of SimulationProgram
    execution (analog_tjm.py:206-279); ``operator`` may be one MPO per interval.
    The noise model is used as given (``Simulator.run`` samples static disorder BEFORE dispatch, ``simulator.py:1269-1271``)."""
    if not isinstance(sim_params, AnalogSimParams):
        raise TypeError("run_batch drives the analog TJM; use digital_tjm / Simulator.run for DigitalSimParams")
    K = _kernels(kernels)
    dev = K.device
    idx = [int(t) for t in traj_indices]
    tensors = _tensors(initial_state)
    length = len(tensors)
    obs = engine_observables(sim_params.sorted_observables, length)
    ep = EngineParams(dt=sim_params.dt, n_steps=len(sim_params.times) - 1, order=2 if sim_params.order != 1 else 1,
                      max_bond_dim=sim_params.max_bond_dim, svd_threshold=sim_params.svd_threshold, trunc_mode=sim_params.trunc_mode,
                      krylov_tol=sim_params.krylov_tol, tdvp_mode=sim_params.tdvp_mode, tdvp_sweeps=sim_params.tdvp_sweeps,
                      sample_timesteps=sim_params.sample_timesteps)
    procs = noise_model.processes if noise_model is not None else None
    sched = noise_model.scheduled_jumps if noise_model is not None else None
    mpo0, per_interval = _interval_mpos(operator, ep.n_steps)
    eng = BatchedTJM(K, length, len(idx), mpo0, ep, procs, obs, tensors, scheduled_jumps=sched, extra_mpos=per_interval or ())
    if rng_mode == "philox":
        tu, su = philox_tables(K, sim_params.random_seed, idx, ep.n_times, ep.order, dev, sample_offset=sample_timestep_offset)
    elif rng_mode == "numpy":
        tu, su = numpy_tables(sim_params.random_seed, idx, ep.n_times, ep.order, sample_offset=sample_timestep_offset)
        if trajectory_rngs is not None:
            if len(trajectory_rngs) != len(idx):
                raise ValueError("trajectory_rngs must hold one Generator per trajectory")
            tu = np.empty((len(idx), 2 * (ep.n_times + 1)))
            rewind = [g.bit_generator.state for g in trajectory_rngs]
            for row, g in enumerate(trajectory_rngs):
                tu[row] = _generator_table(g, ep.n_times + 1)
        tu = torch.from_numpy(tu).to(dev)
        su = torch.from_numpy(su).to(dev) if su is not None else None
    else:
        raise ValueError("rng_mode must be 'numpy' or 'philox'")
    results, diag = eng.run(tu, su, None, sample_at=sample_at, continue_trajectory=continue_trajectory, interval_mpos=per_interval,
                            final_sample_from_trajectory_stream=use_trajectory_rng_for_final_sample and trajectory_rngs is not None)
    status = eng.status.cpu().numpy()
    res_host = results.cpu().numpy()
    if eng.spectra:                                # schmidt_spectrum rows: an object array whose entries are the 500-vectors (analog_tjm.py:424-426)
        obj = np.empty(res_host.shape, dtype=object)
        obj[...] = res_host
        for (k, col), spec in eng.spectra.items():
            sh = spec.cpu().numpy()
            for b in range(len(idx)):
                obj[b, k, col] = sh[b].copy()
        res_host = obj
    if rng_mode == "numpy" and trajectory_rngs is not None:
        for g, state0, used in zip(trajectory_rngs, rewind, eng._pos.cpu().tolist()):
            g.bit_generator.state = state0
            if used:
                g.random(int(used))
    final = None
    if return_states:
        st = eng.psi if ep.order == 2 and ep.n_times > 1 and not return_trajectory_state else eng.phi
        final = [st.download(b) for b in range(len(idx))]
    return res_host, diag.cpu().numpy(), status, final


def _single(args, order: int, *, copy_initial_state, rng, sample_timestep_offset, use_trajectory_rng_for_final_sample, return_trajectory_state,
            continue_trajectory, sample_at, kernels):
    traj_idx, initial_state, noise_model, sim_params, operator = args
    del copy_initial_state                         # the batched engine uploads a copy; the caller's MPS is never mutated
    if order == 1 and (use_trajectory_rng_for_final_sample or return_trajectory_state or continue_trajectory or sample_timestep_offset):
        raise TypeError("analog_tjm_1 takes copy_initial_state, rng and sample_at only (analog_tjm.py:369-376)")
    if sample_at is not None and len(sim_params.times) - 1 not in frozenset(sample_at) and sim_params.get_state and not return_trajectory_state:
        raise NotImplementedError("get_state with a sample_at that omits the final time (the unrecorded physical copy of analog_tjm.py:"
                                  "293-309) is not on the B200 path")
    if order == 2 and any(o.gate.name == "schmidt_spectrum" for o in sim_params.observables):
        # analog_tjm_2 allocates a float result buffer (analog_tjm.py:289-292), so storing the 500-vector fails inside numpy with
        # exactly this message; only the first-order backend (dtype=object, :424-426) can return a Schmidt spectrum
        raise ValueError("setting an array element with a sequence.")
    want = 2 if sim_params.order != 1 else 1
    if want != order:
        raise ValueError(f"analog_tjm_{order} called with AnalogSimParams(order={sim_params.order})")
    res, diag, status, final = run_batch([traj_idx], initial_state, noise_model, sim_params, operator, kernels=kernels,
                                         trajectory_rngs=None if rng is None else [rng], sample_timestep_offset=sample_timestep_offset,
                                         return_states=bool(sim_params.get_state) or return_trajectory_state, sample_at=sample_at,
                                         continue_trajectory=continue_trajectory, return_trajectory_state=return_trajectory_state,
                                         use_trajectory_rng_for_final_sample=use_trajectory_rng_for_final_sample)
    if int(status[0]) != 0:                      # the exception type / message the reference raises in that situation (simulator.STATUS_ERRORS)
        from .simulator import status_exception
        raise status_exception(np.asarray(status, dtype=np.int32))
    final_mps = None
    if final is not None:
        final_mps = MPS(len(final[0]), tensors=final[0])
    return res[0], diag[0], final_mps


def analog_tjm_2(args, *, copy_initial_state: bool = True, rng: np.random.Generator | None = None, sample_timestep_offset: int = 0,
                 use_trajectory_rng_for_final_sample: bool = False, return_trajectory_state: bool = False, continue_trajectory: bool = False,
                 sample_at: Sequence[int] | None = None, kernels=None):
    """analog/analog_tjm.py:206-366 for one trajectory: ``args = (traj_idx, MPS | State, NoiseModel | None, AnalogSimParams,
    MPO | Hamiltonian)`` -> ``(results (n_sorted_obs, T | 1), diagnostics (3, T | 1), final MPS | None)``."""
    return _single(args, 2, copy_initial_state=copy_initial_state, rng=rng, sample_timestep_offset=sample_timestep_offset,
                   use_trajectory_rng_for_final_sample=use_trajectory_rng_for_final_sample, return_trajectory_state=return_trajectory_state,
                   continue_trajectory=continue_trajectory, sample_at=sample_at, kernels=kernels)


def analog_tjm_1(args, *, copy_initial_state: bool = True, rng: np.random.Generator | None = None, sample_timestep_offset: int = 0,
                 use_trajectory_rng_for_final_sample: bool = False, return_trajectory_state: bool = False, continue_trajectory: bool = False,
                 sample_at: Sequence[int] | None = None, kernels=None):
    """analog/analog_tjm.py:369-462 for one trajectory (first-order scheme; scheduled jumps live here)."""
    return _single(args, 1, copy_initial_state=copy_initial_state, rng=rng, sample_timestep_offset=sample_timestep_offset,
                   use_trajectory_rng_for_final_sample=use_trajectory_rng_for_final_sample, return_trajectory_state=return_trajectory_state,
                   continue_trajectory=continue_trajectory, sample_at=sample_at, kernels=kernels)


def digital_tjm(args, *, compiled_circuit=None, kernels=None):
    """digital/digital_tjm.py:636-749 for one trajectory: ``args = (traj_idx, MPS | State, NoiseModel | None, DigitalSimParams,
    circuit)`` -> ``(results | None, diagnostics | None, counts | None, final MPS | None)``.  ``compiled_circuit`` takes
    precedence over the circuit in ``args`` (the reference compiles once per run and passes the layers to every worker)."""
    from .circuits import CompiledCircuit, compile_circuit
    from .digital import plan_digital_shots, shot_tables, validate_digital_noise
    traj_idx, initial_state, noise_model, sp, circuit = args
    if not isinstance(sp, DigitalSimParams):
        raise TypeError("digital_tjm needs DigitalSimParams")
    K = _kernels(kernels)
    dev = K.device
    compiled = compiled_circuit if compiled_circuit is not None else (circuit if isinstance(circuit, CompiledCircuit) else compile_circuit(circuit))
    tensors = _tensors(initial_state)
    length = len(tensors)
    if noise_model is not None:
        validate_digital_noise(noise_model, length)
    wants_obs, wants_shots = bool(sp.observables), sp.shots is not None
    obs = engine_observables(sp.sorted_observables, length)
    ep = EngineParams(dt=1.0, n_steps=1, order=1, max_bond_dim=sp.max_bond_dim, svd_threshold=sp.svd_threshold, trunc_mode=sp.trunc_mode,
                      krylov_tol=sp.krylov_tol, tdvp_mode=sp.tdvp_mode, tdvp_sweeps=sp.tdvp_sweeps, sample_timesteps=True, digital=True)
    mpo = [np.eye(2, dtype=np.complex128).reshape(2, 2, 1, 1) for _ in range(length)]
    procs = list(noise_model.processes) if noise_model is not None else None
    noisy = procs is not None and any(pr["strength"] != 0 for pr in procs)
    eng = BatchedTJM(K, length, 1, mpo, ep, procs, obs, tensors, bond_factor=BatchedTJM.mpo_route_bond_factor(compiled.layers, sp.gate_mode, length))
    n2q = sum(len(l.even_two_qubit_gates) + len(l.odd_two_qubit_gates) for l in compiled.layers)
    tu, _ = numpy_tables(sp.random_seed, [int(traj_idx)], max(n2q, 1), 1)
    res, diag = eng.run_digital(compiled.layers, torch.from_numpy(tu).to(dev), sp.sample_layers, sp.gate_mode, None)
    if int(eng.status.max().item()) != 0:
        from .simulator import status_exception
        raise status_exception(eng.status.cpu().numpy().astype(np.int32))
    counts = None
    if wants_shots:
        _, shots_of = plan_digital_shots(sp, noisy=noisy)
        per = shots_of(int(traj_idx))
        counts = {}
        if per > 0:
            su = shot_tables(K, "numpy", sp.random_seed, [int(traj_idx)], per, length, dev)
            vals, cnt = np.unique(eng.measure_shots(eng.phi, su).cpu().numpy()[0], return_counts=True)
            counts = {int(v): int(c) for v, c in zip(vals, cnt)}
    final = MPS(length, tensors=eng.phi.download(0)) if sp.get_state else None
    if not wants_obs:                              # shots-only: observables and diagnostics stay None (digital_tjm.py:683-686,728-733)
        return None, None, counts, final
    return res.cpu().numpy()[0], diag.cpu().numpy()[0], counts, final
