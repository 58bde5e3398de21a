"""``SimulationProgram``: ordered analog and digital segments that share one evolving state, run as GPU batches.

Reference: the public specification and its compiler (``core/data_structures/simulation_program.py:199-842``) and the per-trajectory
executor (``simulator.py:242-718``, ``_run_program`` ``:1317-1439``).  The reference walks ONE trajectory through all segments in a
forked worker.  Its control flow -- which adjacent analog segments merge into one ``analog_tjm`` call (``:405-427``), where an
order-2 Trotter chain hands its mid-step state ``phi`` to the next segment (``:295-336``), which sample-stream offset each segment
uses -- depends only on the compiled instructions, never on the trajectory.  So the plan is computed once (``plan_units``) and every
unit of it is executed for a whole batch of trajectories in lock step by ``engine.BatchedTJM``:

* the state stays on the device between units (``BatchedState.transfer_from``: a slot-wise device copy into the next unit's layout);
* the single trajectory Generator the reference threads through all segments (``simulator.py:610``) becomes one uniform table per
  trajectory and a device-resident read position that each unit continues (``run(start_pos=...)`` / ``engine._pos``);
* merged analog runs with different Hamiltonians use one MPO per interval (``run(interval_mpos=...)``), zero-padded to a common
  bond dimension.

Piecewise Hamiltonians expand to one MPO per interval exactly like merged runs.  The BUG integrator is not on the B200 path; OpenQASM 2
strings / paths go through the qiskit-free reader (yaqs_b200/qasm.py).
"""
from __future__ import annotations

import copy
from dataclasses import dataclass, field, replace
from typing import Sequence

import numpy as np
import torch

from .circuits import Circuit, CompiledCircuit, compile_circuit
from .containers import Hamiltonian, Result, State
from .engine import BatchedTJM, EngineParams
from .noise import NoiseModel
from .params import AnalogSimParams, DigitalSimParams, Observable, engine_observables
from .rng import make_disorder_rng, numpy_tables, philox_tables

__all__ = ["SimulationProgram"]


@dataclass(frozen=True)
class _AnalogSegment:
    hamiltonian: Hamiltonian
    sim_params: AnalogSimParams
    noise_model: NoiseModel | None = None


@dataclass(frozen=True)
class _DigitalSegment:
    circuit: object
    sim_params: DigitalSimParams
    noise_model: NoiseModel | None = None


def _normalize_program_settings(observables, num_traj, random_seed, get_state):
    """simulation_program.py:71-122."""
    if observables is None:
        obs_tuple: tuple = ()
    else:
        if isinstance(observables, (str, bytes)) or not isinstance(observables, Sequence):
            raise TypeError("observables must be a sequence of Observable.")
        obs_list = list(observables)
        for index, observable in enumerate(obs_list):
            if not isinstance(observable, Observable):
                raise TypeError(f"observables[{index}] must be Observable, got {type(observable).__name__}.")
        obs_tuple = tuple(obs_list)
    if num_traj is not None and (isinstance(num_traj, bool) or not isinstance(num_traj, int)):
        raise TypeError(f"num_traj must be int or None, got {type(num_traj).__name__}.")
    if num_traj is not None and num_traj < 1:
        raise ValueError(f"num_traj must be at least 1, got {num_traj}.")
    if random_seed is not None and (isinstance(random_seed, bool) or not isinstance(random_seed, int)):
        raise TypeError(f"random_seed must be int or None, got {type(random_seed).__name__}.")
    if random_seed is not None and random_seed < 0:
        raise ValueError(f"random_seed must be non-negative, got {random_seed}.")
    if not isinstance(get_state, bool):
        raise TypeError(f"get_state must be bool, got {type(get_state).__name__}.")
    return obs_tuple, num_traj, random_seed, get_state


def _reject_program_owned_fields(sim_params, index: int) -> None:
    """simulation_program.py:125-148."""
    if sim_params.observables:
        raise ValueError(f"segments[{index}] sim_params.observables must be empty; pass observables=... on SimulationProgram instead.")
    if sim_params.random_seed is not None:
        raise ValueError(f"segments[{index}] sim_params.random_seed must be None; pass random_seed=... on SimulationProgram instead.")
    if sim_params.get_state:
        raise ValueError(f"segments[{index}] sim_params.get_state must be False; pass get_state=... on SimulationProgram instead.")


def _normalize_segment(item, index: int):
    """simulation_program.py:151-195."""
    if not isinstance(item, tuple) or len(item) not in {2, 3}:
        raise TypeError(f"segments[{index}] must be a (operator, params) or (operator, params, noise_model) tuple, got {type(item).__name__}.")
    operator, params = item[0], item[1]
    noise_model = item[2] if len(item) == 3 else None
    if noise_model is not None and not isinstance(noise_model, NoiseModel):
        raise TypeError(f"segments[{index}] noise_model must be NoiseModel or None, got {type(noise_model).__name__}.")
    if isinstance(operator, Hamiltonian):
        if not isinstance(params, AnalogSimParams):
            raise TypeError(f"segments[{index}] pairs a Hamiltonian with {type(params).__name__}; expected AnalogSimParams.")
        _reject_program_owned_fields(params, index)
        return _AnalogSegment(operator, params, noise_model)
    if isinstance(operator, str) or hasattr(operator, "__fspath__"):
        from .qasm import load_circuit
        operator = load_circuit(operator)                # OpenQASM 2 text or a path to it (simulation_program.py:181-182)
    if isinstance(operator, (Circuit, CompiledCircuit)) or type(operator).__name__ == "QuantumCircuit":
        if not isinstance(params, DigitalSimParams):
            raise TypeError(f"segments[{index}] pairs a circuit operator with {type(params).__name__}; expected DigitalSimParams.")
        _reject_program_owned_fields(params, index)
        return _DigitalSegment(operator, params, noise_model)
    raise TypeError(f"segments[{index}] operator must be Hamiltonian, QuantumCircuit, or an OpenQASM str/Path, got {type(operator).__name__}.")


class SimulationProgram:
    """simulation_program.py:199-293: immutable ordered program; observables, seed and ``get_state`` are program-wide."""

    __slots__ = ("segments", "observables", "num_traj", "random_seed", "get_state")

    def __init__(self, segments, *, observables=None, num_traj: int | None = None, random_seed: int | None = None, get_state: bool = False) -> None:
        if isinstance(segments, (str, bytes)):
            raise TypeError("segments must be an iterable of (operator, params) pairs.")
        try:
            raw = tuple(segments)
        except TypeError as error:
            raise TypeError("segments must be an iterable of (operator, params) pairs.") from error
        if not raw:
            raise ValueError("SimulationProgram requires at least one segment.")
        normalized = tuple(_normalize_segment(item, index) for index, item in enumerate(raw))
        obs_tuple, num_traj, random_seed, get_state = _normalize_program_settings(observables, num_traj, random_seed, get_state)
        for name, value in (("segments", normalized), ("observables", obs_tuple), ("num_traj", num_traj), ("random_seed", random_seed),
                            ("get_state", get_state)):
            object.__setattr__(self, name, value)

    def __setattr__(self, name, value):
        raise AttributeError("SimulationProgram is immutable")

    def __iter__(self):
        return iter(self.segments)

    def __len__(self) -> int:
        return len(self.segments)


# ----------------------------------------------------------------------------------------------------------------------
# compilation (simulation_program.py:306-755)
# ----------------------------------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class _AnalogInstruction:
    index: int
    hamiltonian: object                 # containers.MPO, or a tuple with one MPO per interval (piecewise Hamiltonian)
    sim_params: AnalogSimParams
    execution_params: AnalogSimParams
    noise_model: NoiseModel | None
    time_offset: float


@dataclass(frozen=True)
class _DigitalInstruction:
    index: int
    circuit: object
    compiled_circuit: CompiledCircuit
    sim_params: DigitalSimParams
    execution_params: DigitalSimParams
    noise_model: NoiseModel | None
    time_offset: float


@dataclass(frozen=True)
class _CompiledProgram:
    instructions: tuple
    length: int
    observables: tuple
    get_state: bool
    num_traj: int
    random_seed: int | None
    default_noise_model: NoiseModel | None


def _validate_noise_layout(noise_model, length: int, segment_index: int) -> None:
    """simulation_program.py:346-390 for qubit chains."""
    if noise_model is None:
        return
    for entry in [*noise_model.processes, *noise_model.scheduled_jumps]:
        sites = tuple(entry["sites"])
        if not sites or any(not isinstance(s, (int, np.integer)) or s < 0 or s >= length for s in sites):
            raise ValueError(f"segments[{segment_index}] noise process has invalid target sites {list(sites)}.")
        factors = entry.get("factors")
        if factors is not None:
            if len(factors) != len(sites):
                raise ValueError(f"segments[{segment_index}] noise process has {len(factors)} factors for {len(sites)} sites.")
            for site, factor in zip(sites, factors):
                if np.shape(factor) != (2, 2):
                    raise ValueError(f"segments[{segment_index}] noise operator on site {site} has shape {np.shape(factor)}, expected (2, 2).")
            continue
        matrix = entry.get("matrix")
        if matrix is not None and np.shape(matrix) != (2 ** len(sites),) * 2:
            raise ValueError(f"segments[{segment_index}] noise operator on sites {list(sites)} has shape {np.shape(matrix)}, "
                             f"expected {(2 ** len(sites),) * 2}.")


def _resolve_program_num_traj(program: SimulationProgram) -> int:
    """simulation_program.py:393-413."""
    if program.num_traj is not None:
        return program.num_traj
    counts = {segment.sim_params.num_traj for segment in program.segments}
    if len(counts) == 1:
        return next(iter(counts))
    raise ValueError("segments disagree on sim_params.num_traj; pass num_traj=... on SimulationProgram to select the ensemble size.")


def _apply_program_settings(sim_params, observables, num_traj: int, random_seed):
    """simulation_program.py:416-440: execution copies always hand their state on (``get_state=True``)."""
    execution_params = copy.deepcopy(sim_params)
    execution_params.get_state = True
    execution_params.observables = [copy.deepcopy(o) for o in observables]
    execution_params.num_traj = num_traj
    execution_params.random_seed = random_seed
    return execution_params


def _duration_to_steps(duration: float, dt: float, label: str) -> int:
    """simulation_program.py:443-470."""
    n_steps = round(duration / dt)
    evolved = n_steps * dt
    residual = abs(duration - evolved)
    roundoff_tol = max(np.spacing(duration), abs(n_steps) * np.spacing(dt), 8 * np.finfo(np.float64).eps * max(duration, evolved, dt))
    if n_steps <= 0 or residual > min(roundoff_tol, 0.25 * dt):
        raise ValueError(f"{label} ({duration}) must be an integer multiple of dt ({dt}).")
    return n_steps


def expand_analog_operator(hamiltonian: Hamiltonian, sim_params: AnalogSimParams, label: str = ""):
    """_expand_analog_operator (simulation_program.py:490-547): a static Hamiltonian gives its MPO, a piecewise one a tuple with
    one MPO reference per analog interval (collapsed to the bare MPO when every interval shares the same object)."""
    prefix = f"{label} " if label else ""
    if not hamiltonian.is_piecewise:
        return hamiltonian.ensure_mpo().mpo
    dt = float(sim_params.dt)
    n_intervals = max(len(sim_params.times) - 1, 0)
    interval_mpos: list = []
    for index, (piece, duration) in enumerate(hamiltonian.pieces):
        n_steps = _duration_to_steps(duration, dt, f"{prefix}pieces[{index}] duration")
        interval_mpos.extend([piece.ensure_mpo().mpo] * n_steps)
    if len(interval_mpos) != n_intervals:
        raise ValueError(f"{prefix}piecewise durations must sum to elapsed_time ({sim_params.elapsed_time}); got {len(interval_mpos) * dt}, "
                         f"expected {sim_params.elapsed_time}.")
    if interval_mpos and all(m is interval_mpos[0] for m in interval_mpos):
        return interval_mpos[0]
    return tuple(interval_mpos)


def compile_program(program: SimulationProgram, initial_state: State, default_noise_model: NoiseModel | None = None) -> _CompiledProgram:
    """simulation_program.py:679-755."""
    if initial_state.representation != "mps":
        raise ValueError("SimulationProgram execution currently requires State.representation='mps'.")
    length = initial_state.length
    num_traj = _resolve_program_num_traj(program)
    instructions: list = []
    time_offset = 0.0
    for index, segment in enumerate(program.segments):
        noise = segment.noise_model if segment.noise_model is not None else default_noise_model
        _validate_noise_layout(noise, length, index)
        params = _apply_program_settings(segment.sim_params, program.observables, num_traj, program.random_seed)
        if isinstance(segment, _AnalogSegment):
            if segment.hamiltonian.length != length:
                raise ValueError(f"segments[{index}] Hamiltonian.length={segment.hamiltonian.length} does not match State.length={length}.")
            if segment.sim_params.order not in {1, 2}:
                raise ValueError(f"segments[{index}] AnalogSimParams.order must be 1 or 2, got {segment.sim_params.order}.")
            if segment.sim_params.multi_time_observables:
                raise ValueError(f"segments[{index}] multi_time_observables are not supported in program execution.")
            operator = expand_analog_operator(segment.hamiltonian, params, label=f"segments[{index}]")
            instructions.append(_AnalogInstruction(index, operator, segment.sim_params, params, noise, time_offset))
            time_offset += float(segment.sim_params.elapsed_time)
            continue
        circuit = segment.circuit
        compiled = circuit if isinstance(circuit, CompiledCircuit) else compile_circuit(circuit)
        nq = getattr(circuit, "num_qubits", length)
        if nq != length:
            raise ValueError(f"segments[{index}] circuit.num_qubits={nq} does not match State.length={length}.")
        if params.sample_layers:
            params.num_mid_measurements = compiled.num_mid_measurements
        instructions.append(_DigitalInstruction(index, circuit, compiled, segment.sim_params, params, noise, time_offset))
    return _CompiledProgram(tuple(instructions), length, program.observables, program.get_state, num_traj, program.random_seed, default_noise_model)


def _is_stochastic(noise_model) -> bool:
    """simulator.py:242-246."""
    return noise_model is not None and any(float(p["strength"]) > 0 for p in noise_model.processes)


def sample_program_noise_models(compiled: _CompiledProgram) -> _CompiledProgram:
    """simulator.py:249-274: each distinct model object is sampled once per run, in order of first use (default first)."""
    rng = make_disorder_rng(compiled.random_seed)
    sampled: dict[int, NoiseModel] = {}

    def sample(nm):
        if nm is None:
            return None
        if id(nm) not in sampled:
            sampled[id(nm)] = nm.sample(rng=rng)
        return sampled[id(nm)]

    default = sample(compiled.default_noise_model)
    return replace(compiled, instructions=tuple(replace(i, noise_model=sample(i.noise_model)) for i in compiled.instructions),
                   default_noise_model=default)


def validate_compiled_program(compiled: _CompiledProgram):
    """simulator.py:721-762 -> (has_observables, has_shots, stochastic)."""
    from .digital import validate_digital_noise
    from .simulator import validate_noise_model_for_run
    has_obs = bool(compiled.observables)
    has_shots = any(isinstance(i, _DigitalInstruction) and i.execution_params.shots is not None for i in compiled.instructions)
    if not has_obs and not has_shots and not compiled.get_state:
        raise ValueError("No output specified: set observables=, shots on a digital segment, and/or get_state=.")
    for ins in compiled.instructions:
        if ins.noise_model is None:
            continue
        if isinstance(ins, _DigitalInstruction):
            validate_digital_noise(ins.noise_model, compiled.length)
        else:
            validate_noise_model_for_run(ins.noise_model, compiled.length, ins.execution_params)
    stochastic = any(_is_stochastic(i.noise_model) for i in compiled.instructions)
    if stochastic and compiled.get_state:
        raise ValueError("Cannot return state from a noisy SimulationProgram due to stochastic trajectories.")
    for ins in compiled.instructions:
        if ins.sim_params.get_state and _is_stochastic(ins.noise_model):
            raise ValueError("Cannot return state from a noisy SimulationProgram due to stochastic trajectories.")
    return has_obs, has_shots, stochastic


# ----------------------------------------------------------------------------------------------------------------------
# the execution plan: the trajectory-independent control flow of _execute_program_trajectory (simulator.py:588-718)
# ----------------------------------------------------------------------------------------------------------------------
def _same_order2_operator(left: _AnalogInstruction, right: _AnalogInstruction) -> bool:
    """simulator.py:277-292."""
    if left.noise_model is not right.noise_model:
        return False
    lm, rm = left.hamiltonian, right.hamiltonian
    if isinstance(lm, tuple) or isinstance(rm, tuple):
        return False
    if lm is rm:
        return True
    if len(lm.tensors) != len(rm.tensors):
        return False
    return all(np.array_equal(a, b) for a, b in zip(lm.tensors, rm.tensors))


def _order2_chain_continues(instructions, current: _AnalogInstruction, index: int) -> bool:
    """simulator.py:295-320."""
    nxt = instructions[index + 1] if index + 1 < len(instructions) else None
    if not isinstance(nxt, _AnalogInstruction):
        return False
    cp, np_ = current.execution_params, nxt.execution_params
    return (np_.order == 2 and len(cp.times) > 1 and len(np_.times) > 1 and cp.dt == np_.dt and cp.evolution_mode == np_.evolution_mode
            and cp.sample_timesteps == np_.sample_timesteps and _same_order2_operator(current, nxt))


def _order2_hand_off(instructions, current: _AnalogInstruction) -> bool:
    """simulator.py:323-336."""
    if not _order2_chain_continues(instructions, current, current.index):
        return False
    nxt = instructions[current.index + 1]
    return not current.sim_params.get_state and not nxt.sim_params.get_state


def _analog_runs_can_merge(left: _AnalogInstruction, right: _AnalogInstruction) -> bool:
    """simulator.py:405-426."""
    lp, rp = left.execution_params, right.execution_params
    if left.sim_params.get_state or right.sim_params.get_state:
        return False
    return (lp.order == rp.order and lp.dt == rp.dt and lp.sample_timesteps == rp.sample_timesteps and left.noise_model is right.noise_model
            and len(lp.times) > 1 and len(rp.times) > 1 and lp.max_bond_dim == rp.max_bond_dim and lp.svd_threshold == rp.svd_threshold
            and lp.trunc_mode == rp.trunc_mode and lp.krylov_tol == rp.krylov_tol and lp.tdvp_sweeps == rp.tdvp_sweeps
            and lp.tdvp_mode == rp.tdvp_mode)


@dataclass
class _Unit:
    """One batched engine invocation: a digital instruction, a single analog instruction, or a merged analog run."""
    kind: str                                   # "digital" | "analog"
    instructions: list
    sample_offset: int = 0                      # order 2: global index of the first measurement stream (simulator.py:378-383)
    continue_trajectory: bool = False           # order 2: the incoming state is a mid-Trotter phi
    hand_off: bool = False                      # order 2: pass phi (not the last measurement copy) to the next unit
    n_steps: int = 0
    interval_mpos: list | None = None           # one MPO (list of site tensors) per interval when the merged Hamiltonians differ
    sample_at: tuple | None = None
    columns: list = field(default_factory=list)  # per instruction: (start, end) columns of the unit's buffers


def plan_units(instructions: Sequence) -> list[_Unit]:
    units: list[_Unit] = []
    offset, cont = 0, False
    index = 0
    while index < len(instructions):
        ins = instructions[index]
        if isinstance(ins, _DigitalInstruction):
            units.append(_Unit("digital", [ins]))
            offset, cont = 0, False
            index += 1
            continue
        run = [ins]
        while not cont and index + len(run) < len(instructions):
            nxt = instructions[index + len(run)]
            if not isinstance(nxt, _AnalogInstruction) or not _analog_runs_can_merge(run[-1], nxt):
                break
            run.append(nxt)
        steps = [max(len(i.execution_params.times) - 1, 0) for i in run]
        order2 = run[0].execution_params.order != 1
        # the merged path asks for the hand-off whatever the run's own order is (simulator.py:677-692); a single order-1
        # instruction never does (:392-403)
        hand = _order2_hand_off(instructions, run[-1]) if (order2 or len(run) > 1) else False
        unit = _Unit("analog", run, n_steps=sum(steps), hand_off=hand)
        mpos = [m for i, n in zip(run, steps) for m in (i.hamiltonian if isinstance(i.hamiltonian, tuple) else (i.hamiltonian,) * n)]
        if mpos and not all(m is mpos[0] for m in mpos):
            unit.interval_mpos = [m.tensors for m in mpos]
        if len(run) == 1:
            unit.sample_offset, unit.continue_trajectory = (offset, cont) if order2 else (0, False)
            ncol = len(ins.execution_params.times) if ins.execution_params.sample_timesteps else 1
            unit.columns = [(0, ncol)]
            offset = offset + max(len(ins.execution_params.times) - 1, 0) if hand else 0
        else:                                   # merged run: a fresh Trotter chain on the concatenated grid (simulator.py:514-551)
            user_sample = run[0].execution_params.sample_timesteps
            bounds = list(np.cumsum(steps))
            unit.sample_at = None if user_sample else tuple(int(b) for b in bounds)
            start = 0
            for n in steps:                     # _split_merged_analog_results (simulator.py:485-511)
                unit.columns.append((start, start + n + 1) if user_sample else (start + n, start + n + 1))
                start += n
            offset = sum(steps) if hand else 0
        cont = hand
        units.append(unit)
        index += len(run)
    return units


# ----------------------------------------------------------------------------------------------------------------------
# batched execution (simulator.py:1317-1439 with the fork pool replaced by lock-step batches)
# ----------------------------------------------------------------------------------------------------------------------
def _engine_params(sp, n_steps: int, sample_timesteps: bool) -> EngineParams:
    if isinstance(sp, DigitalSimParams):
        return EngineParams(dt=1.0, n_steps=1, order=1, max_bond_dim=sp.max_bond_dim, svd_threshold=sp.svd_threshold, trunc_mode=sp.trunc_mode,
                            krylov_tol=sp.krylov_tol, tdvp_mode=sp.tdvp_mode, tdvp_sweeps=sp.tdvp_sweeps, sample_timesteps=True, digital=True)
    return EngineParams(dt=sp.dt, n_steps=n_steps, order=2 if sp.order != 1 else 1, max_bond_dim=sp.max_bond_dim, svd_threshold=sp.svd_threshold,
                        trunc_mode=sp.trunc_mode, krylov_tol=sp.krylov_tol, tdvp_mode=sp.tdvp_mode, tdvp_sweeps=sp.tdvp_sweeps,
                        sample_timesteps=sample_timesteps)


def _first_mpo(ins: _AnalogInstruction):
    return ins.hamiltonian[0] if isinstance(ins.hamiltonian, tuple) else ins.hamiltonian


def _n_two_qubit(compiled: CompiledCircuit) -> int:
    return sum(len(l.even_two_qubit_gates) + len(l.odd_two_qubit_gates) for l in compiled.layers)


def run_program(simulator, initial_state: State, program: SimulationProgram, noise_model: NoiseModel | None) -> Result:
    from .digital import aggregate_counts, shot_tables
    from .simulator import shard_range
    compiled = sample_program_noise_models(compile_program(program, initial_state, noise_model))
    has_obs, _has_shots, stochastic = validate_compiled_program(compiled)
    if stochastic:
        budgets = [i.execution_params.shots for i in compiled.instructions if isinstance(i, _DigitalInstruction) and i.execution_params.shots is not None]
        ntraj = max(budgets) if budgets and not has_obs else compiled.num_traj
    else:
        ntraj = 1
    length = compiled.length
    units = plan_units(compiled.instructions)
    sorted_ref = compiled.instructions[0].execution_params
    obs = engine_observables(sorted_ref.sorted_observables, length)
    to_sorted = sorted_ref.observable_sorted_indices
    nobs = len(obs)

    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    mine = shard_range(ntraj, rank, world)
    K = simulator._get_kernels()
    dev = K.device

    def ncols(ins) -> int:
        sp = ins.execution_params
        if isinstance(ins, _DigitalInstruction):
            return (ins.compiled_circuit.num_mid_measurements + 2) if sp.sample_layers else 1
        return len(sp.times) if sp.sample_timesteps else 1

    res_local = [np.zeros((len(mine), nobs, ncols(i))) for i in compiled.instructions]
    diag_local = [np.zeros((len(mine), 3, ncols(i))) for i in compiled.instructions]
    status_local = np.zeros(len(mine), dtype=np.int32)
    draws_local = np.zeros(len(mine), dtype=np.int64)      # doubles each trajectory's stream handed out (jump tests + selections)
    shots_local: list[dict[int, dict[int, int]]] = [{} for _ in compiled.instructions]
    checkpoints: dict[int, list] = {}
    final_tensors = None
    stats: dict = {}
    n_tests = 1 + sum(u.n_steps + 1 if u.kind == "analog" else _n_two_qubit(u.instructions[0].compiled_circuit) for u in units)
    ident = [np.eye(2, dtype=np.complex128).reshape(2, 2, 1, 1) for _ in range(length)]
    def sp_layers(unit):
        return unit.instructions[0].compiled_circuit.layers

    # waves: the most demanding unit decides how many trajectories fit (states of two engines coexist only during a hand-over)
    wave = len(mine)
    for u in units:
        sp = u.instructions[0].execution_params
        ep = _engine_params(sp, u.n_steps, True)
        mpo = ident if u.kind == "digital" else _first_mpo(u.instructions[0]).tensors
        factor = BatchedTJM.mpo_route_bond_factor(sp_layers(u), sp.gate_mode, length) if u.kind == "digital" else None
        wave = min(wave, simulator._wave_size(len(mine), length, ep, mpo, initial_state.mps.tensors, dev, bond_factor=factor))
    wave = max(wave, 1)
    K.launch_count(reset=True)
    for w0 in range(0, len(mine), wave):
        idx = list(mine[w0: w0 + wave])
        rows = slice(w0, w0 + len(idx))
        if simulator.rng_mode == "numpy":
            tu = torch.from_numpy(numpy_tables(compiled.random_seed, idx, n_tests, 1)[0]).to(dev)
        else:
            tu = philox_tables(K, compiled.random_seed, idx, n_tests, 1, dev)[0]
        pos = torch.zeros(len(idx), dtype=torch.int64, device=dev)
        carried = initial_state.mps.tensors          # host MPS (broadcast) for the first unit, BatchedState afterwards
        eng = None
        for u in units:
            first = u.instructions[0]
            sp = first.execution_params
            procs = first.noise_model.processes if first.noise_model is not None else None
            if u.kind == "digital":
                ep = _engine_params(sp, 1, True)
                new = BatchedTJM(K, length, len(idx), ident, ep, list(procs) if procs is not None else None, obs, carried,
                                 bond_factor=BatchedTJM.mpo_route_bond_factor(first.compiled_circuit.layers, sp.gate_mode, length))
                carried = eng = None                 # the constructor has copied the state: release the previous unit's engine NOW
                r, dg = new.run_digital(first.compiled_circuit.layers, tu, sp.sample_layers, sp.gate_mode, None, start_pos=pos)
                out_state = new.phi
                if sp.shots is not None:             # _per_call_shots with shot_distribution = (shots, ntraj) (digital_tjm.py:752-768)
                    base, rem = divmod(int(sp.shots), int(ntraj))
                    per = np.array([base + (1 if t < rem else 0) for t in idx], dtype=np.int64)
                    if per.max() > 0:
                        su = shot_tables(K, simulator.rng_mode, compiled.random_seed, idx, int(per.max()), length, dev)
                        outs = new.measure_shots(new.phi, su).cpu().numpy()
                        for row, t in enumerate(idx):
                            vals, cnt = np.unique(outs[row, : per[row]], return_counts=True)
                            shots_local[first.index][t] = {int(v): int(c) for v, c in zip(vals, cnt)} if per[row] > 0 else {}
                    else:
                        for t in idx:
                            shots_local[first.index][t] = {}
                res_local[first.index][rows] = r.cpu().numpy()
                diag_local[first.index][rows] = dg.cpu().numpy()
            else:
                merged = len(u.instructions) > 1
                ep = _engine_params(sp, u.n_steps, True if merged else sp.sample_timesteps)
                sched = first.noise_model.scheduled_jumps if first.noise_model is not None else None
                extra = u.interval_mpos or ()
                new = BatchedTJM(K, length, len(idx), _first_mpo(first).tensors, ep, procs, obs, carried, scheduled_jumps=sched, extra_mpos=extra)
                carried = eng = None                 # (waves are sized for ONE engine; two alive while a unit runs would not fit)
                su = None
                if ep.order == 2:
                    if simulator.rng_mode == "numpy":
                        su = torch.from_numpy(numpy_tables(compiled.random_seed, idx, ep.n_times, 2, sample_offset=u.sample_offset)[1]).to(dev)
                    else:
                        su = philox_tables(K, compiled.random_seed, idx, ep.n_times, 2, dev, sample_offset=u.sample_offset)[1]
                r, dg = new.run(tu, su, None, start_pos=pos, sample_at=u.sample_at, continue_trajectory=u.continue_trajectory,
                                interval_mpos=u.interval_mpos)
                # what the next unit starts from: phi on an order-2 hand-off, else the physical state at the final time
                # (the last measurement copy for order 2, analog_tjm.py:357-366; the evolved state itself for order 1, :460)
                out_state = new.phi if (ep.order == 1 or u.hand_off or ep.n_times == 1) else new.psi
                r, dg = r.cpu().numpy(), dg.cpu().numpy()
                for ins, (c0, c1) in zip(u.instructions, u.columns):
                    res_local[ins.index][rows] = r[:, :, c0:c1]
                    diag_local[ins.index][rows] = dg[:, :, c0:c1]
            pos = new._pos
            status_local[rows] |= new.status.cpu().numpy()
            for k, v in new.stats.items():
                stats[k] = stats.get(k, 0) + v
            last = u.instructions[-1]
            if last.sim_params.get_state and not _is_stochastic(last.noise_model):
                checkpoints[last.index] = out_state.download(0)
            carried, eng = out_state, new         # alive only until the next unit's constructor has copied the state
        draws_local[rows] = pos.cpu().numpy()
        if compiled.get_state:
            final_tensors = carried.download(0)
        del eng
    stats["kernel_launches"] = K.launch_count()

    segment_results: list[Result] = []
    for ins in compiled.instructions:
        sp = ins.execution_params
        seg = Result(sim_params=sp, noise_model=ins.noise_model)
        seg.segment_index, seg.time_offset = ins.index, ins.time_offset
        seg.segment_type = "analog" if isinstance(ins, _AnalogInstruction) else "digital"
        digital = isinstance(ins, _DigitalInstruction)
        shots_only = digital and sp.shots is not None and not has_obs
        if digital:
            times = None
        else:
            times = np.asarray(sp.times, dtype=np.float64) if sp.sample_timesteps else np.asarray([sp.elapsed_time], dtype=np.float64)
        simulator._finish(seg, sp, res_local[ins.index], diag_local[ins.index], status_local, mine, ntraj, nobs, ncols(ins), to_sorted, times,
                          dict(stats), dev)
        if shots_only:                               # digital_tjm.py:728-733: no observables, no diagnostics
            seg.runtime_cost = seg.max_bond = seg.total_bond = None
        if digital and sp.shots is not None:
            local = shots_local[ins.index]
            if world > 1:
                gathered: list = [None] * world
                dist.all_gather_object(gathered, local)
                local = {k: v for part in gathered for k, v in part.items()}
            seg.measurements = [local.get(t) for t in range(ntraj)]
            aggregate_counts(seg)
        if ins.index in checkpoints:
            seg.output_state = State(tensors=checkpoints[ins.index])
        segment_results.append(seg)

    values, times, counts = stitch_program_results(segment_results, compiled.observables)
    result = Result(observables=[copy.deepcopy(o) for o in compiled.observables], expectation_values=values, times=times, counts=counts,
                    noise_model=compiled.default_noise_model)
    result.segment_results = segment_results
    result.stats = stats
    result.stats["trajectory_draws"] = draws_local        # this rank's shard
    result.status = status_local
    if final_tensors is not None:
        result.output_state = State(tensors=final_tensors)
    return result


def stitch_program_results(segment_results: Sequence[Result], observables: Sequence):
    """simulation_program.py:758-842: concatenate per-segment expectation values on the physical program timeline; the outer
    ``counts`` are those of the last segment that recorded shots."""
    counts = None
    for seg in reversed(segment_results):
        if seg.counts:
            counts = dict(sorted(seg.counts.items()))
            break
    if not observables:
        return [], None, counts
    parts: list[list[np.ndarray]] = [[] for _ in observables]
    time_parts: list[np.ndarray] = []
    for seg in segment_results:
        if len(seg.expectation_values) != len(observables):
            raise ValueError(f"Segment {seg.segment_index} recorded {len(seg.expectation_values)} observables, expected {len(observables)}.")
        n = len(np.asarray(seg.expectation_values[0]))
        offset = 0.0 if seg.time_offset is None else float(seg.time_offset)
        if seg.segment_type == "digital":
            time_parts.append(np.full(n, offset, dtype=np.float64))
        else:
            if seg.times is None:
                raise ValueError(f"Analog segment {seg.segment_index} has no time data.")
            t = np.asarray(seg.times, dtype=np.float64)
            if t.ndim != 1 or len(t) != n:
                raise ValueError(f"Observable in analog segment {seg.segment_index} is not aligned with times.")
            time_parts.append(t + offset)
        for k, v in enumerate(seg.expectation_values):
            arr = np.asarray(v)
            if arr.ndim != 1 or len(arr) != n:
                raise ValueError(f"Segment {seg.segment_index} observable {k} has inconsistent shape.")
            parts[k].append(arr)
    return [np.concatenate(p) for p in parts], (np.concatenate(time_parts) if time_parts else None), counts
