"""``Simulator.run`` for the TJM hot path: same call, same ``Result`` as the reference, trajectories run as GPU batches.

Reference: ``Simulator`` (simulator.py:1059-1111), ``run`` (:1173-1312), ``_run_analog`` (:1444-1682).  The reference fans
trajectories out over a fork pool (core/parallel_utils.py:331-390); here ``run`` slices the trajectory index range into
waves that fit the GPU's HBM, runs every wave as one lock-step batch on the device (``engine.BatchedTJM``) and, when
``torch.distributed`` is initialised with more than one rank, shards the index range over ranks and combines the
per-column sums and sample counts with ONE all-reduce (NCCL over NVLink on GPU ranks, gloo in CPU tests).
"""
from __future__ import annotations

import math
from typing import Sequence

import numpy as np
import torch

from .containers import Hamiltonian, Result, State
from .engine import BatchedTJM, BatchLayout, EngineParams, layout_chi
from .noise import NoiseModel
from .params import AnalogSimParams, DigitalSimParams, engine_observables
from .rng import make_disorder_rng, numpy_tables, philox_source


# Per-trajectory device status bits (engine.STATUS_*) and the exception the reference raises in the same situation:
#   4  zero / non-finite jump weights        ValueError, stochastic_process.py:179-187 (raised where it happens, mid-run)
#   2  non-finite singular values            ValueError, scipy.linalg.svd(check_finite=True) behind core/linalg/svd.py:89-104
#   1  SVD iteration did not converge        numpy.linalg.LinAlgError, the LAPACK failure surfaced by the same call
#   8  observable with an imaginary part     AssertionError, mps.py:1233
STATUS_ERRORS = (
    (4, ValueError, "Jump probability weights are zero or non-finite. Reduce process strengths and/or the timestep dt so that "
                    "dt * strength * ||L|psi>||^2 remains representable."),
    (2, ValueError, "array must not contain infs or NaNs"),
    (1, np.linalg.LinAlgError, "SVD did not converge"),
    (8, AssertionError, "Measurement should be real"),
)


class TrajectoryError(Exception):
    """Mixin carried by the exception ``Simulator.run`` raises for flagged trajectories: ``result`` (the complete Result, rows of
    the flagged trajectories included), ``status`` (int32 per trajectory) and ``trajectories`` (the flagged indices)."""
    result = None
    status = None
    trajectories = ()


def status_exception(status_all: np.ndarray, result=None) -> Exception | None:
    """The reference-typed exception for a status vector (None when it is clean).  The class is created on the fly so that it is
    BOTH the reference's type (``except ValueError`` / ``LinAlgError`` / ``AssertionError`` keep working) and a TrajectoryError."""
    bad = np.nonzero(status_all)[0]
    if bad.size == 0:
        return None
    word = int(np.bitwise_or.reduce(status_all[bad]))
    for bit, etype, msg in STATUS_ERRORS:
        if word & bit:
            cls = type(f"Trajectory{etype.__name__}", (etype, TrajectoryError), {})
            hit = [int(t) for t in bad if int(status_all[t]) & bit]
            exc = cls(f"{msg} (trajectories {hit[:8]}{'...' if len(hit) > 8 else ''}; status words are in .status, the complete "
                      "Result in .result)")
            exc.result, exc.status, exc.trajectories = result, status_all, tuple(int(t) for t in bad)
            return exc
    exc = type("TrajectoryRuntimeError", (RuntimeError, TrajectoryError), {})(f"unknown device status word {word}")
    exc.result, exc.status, exc.trajectories = result, status_all, tuple(int(t) for t in bad)
    return exc


def shard_range(n: int, rank: int, world: int) -> range:
    """Contiguous trajectory-index shard of rank `rank`."""
    per = math.ceil(n / world) if world > 0 else n
    return range(min(n, rank * per), min(n, (rank + 1) * per))


def _check_operator_dims(entry: dict, kind: str, length: int) -> None:
    """noise_model.py:695-722: sites inside the chain, matrix / factor shapes for qubits; the reference's messages."""
    sites = entry["sites"]
    for site in sites:
        if site >= length:
            raise ValueError(f"{kind} site index {site} is out of range for length {length}.")
    if "matrix" in entry:
        shape = np.asarray(entry["matrix"]).shape
        expected = 2 if len(sites) == 1 else 4
        if shape != (expected, expected):
            raise ValueError(f"{kind} matrix shape {shape} does not match expected ({expected}, {expected}) for sites {sites}.")
    if "factors" in entry:
        for site, factor in zip(sites, entry["factors"], strict=True):
            shape = np.asarray(factor).shape
            if shape != (2, 2):
                raise ValueError(f"{kind} factor on site {site} has shape {shape}, expected (2, 2).")


def validate_noise_model_for_run(noise: NoiseModel, length: int, sim_params=None, is_digital: bool = False) -> None:
    """core/data_structures/noise_model.py:668-780, in the reference's order and with its messages: processes first (sites, operator
    shapes, what the digital / analog MPS paths cannot take), then the scheduled jumps."""
    from .noise import is_pauli
    for pr in noise.processes:
        _check_operator_dims(pr, "Process", length)
        long_range = len(pr["sites"]) == 2 and abs(pr["sites"][1] - pr["sites"][0]) != 1
        if is_digital and long_range:
            raise ValueError(f"Digital TJM does not support non-adjacent / factorized two-site noise (process '{pr['name']}' on sites "
                             f"{pr['sites']}). Gate-local digital noise scoping remains nearest-neighbor only.")
        if not is_digital and long_range and not is_pauli(pr):
            raise ValueError(f"Analog MPS TJM does not support non-Pauli long-range noise (process '{pr['name']}' on sites {pr['sites']}).")
    if not noise.scheduled_jumps:
        return
    if is_digital or not isinstance(sim_params, AnalogSimParams):
        raise ValueError("scheduled_jumps are only supported for single-State analog MPS TJM; they are not supported for MCWF, "
                         "Lindblad, digital, or list[State] ensemble runs.")
    if sim_params.order != 1:
        raise ValueError(f"scheduled_jumps are only supported for AnalogSimParams(order=1); got order={sim_params.order}. Order-2 TJM "
                         "applies deterministic jumps inconsistently on the sampling versus trajectory MPS.")
    times = np.asarray(sim_params.times, dtype=float)
    atol = sim_params.dt * 1e-3
    for jump in noise.scheduled_jumps:
        _check_operator_dims(jump, "Scheduled jump", length)
        jump_time = float(jump["time"])
        if not np.any(np.isclose(times, jump_time, atol=atol, rtol=0.0)):
            raise ValueError(f"Scheduled jump time {jump_time} is not on the simulation time grid (atol={atol}, rtol=0).")


def bytes_per_trajectory(lay_one: BatchLayout, order: int, one_site: bool) -> int:
    """Device bytes one trajectory occupies (DESIGN.md 'Data layout in HBM'); used to size waves."""
    c = 16
    state = lay_one.state_elems * (2 if order == 2 else 1)
    env = lay_one.env_elems * (2 if one_site else 1)
    work = lay_one.theta_slot + 2 * lay_one.t_slot + 26 * lay_one.nloc_max + lay_one.nloc_max + 4 * lay_one.svd_dim ** 2
    return c * (state + env + work) + 4096


class Simulator:
    """Drop-in for ``mqt.yaqs.Simulator``.  ``parallel`` / ``max_workers`` / ``mp_context`` / retry options are accepted for
    signature compatibility and ignored: there is no process pool, the batch is the parallelism."""

    def __init__(self, *, parallel: bool = True, max_workers: int | None = None, show_progress: bool = True, mp_context=None,
                 max_retries: int = 10, retry_exceptions=None, device: str | torch.device | None = None, kernels=None,
                 batch_size: int | None = None, rng_mode: str = "numpy", collect_log: bool = False, lanes: int | None = None,
                 on_trajectory_error: str = "raise") -> None:
        self.parallel, self.max_workers, self.show_progress = parallel, max_workers, show_progress
        self.device = device
        self._kernels = kernels
        self.batch_size = batch_size
        if rng_mode not in ("numpy", "philox"):
            raise ValueError("rng_mode must be 'numpy' (reference-identical seeded streams) or 'philox' (device generator)")
        self.rng_mode = rng_mode
        self.collect_log = collect_log
        self.record_step_events = False     # bench.py: keep CUDA events around the device work of every time step (AnalogRun.lane_runs)
        self.last_log: list | None = None
        self.lanes = lanes          # concurrent sub-batches per GPU (None: 2 on CUDA for waves of >= 64 trajectories, else 1)
        if on_trajectory_error not in ("raise", "flag"):
            raise ValueError("on_trajectory_error must be 'raise' (the reference's behaviour: the run fails with the exception of "
                             "the failing trajectory) or 'flag' (return the Result; flagged trajectories are listed in Result.status "
                             "and excluded from the means)")
        self.on_trajectory_error = on_trajectory_error

    # ------------------------------------------------------------------------------------------------------------------
    def _get_kernels(self):
        if self._kernels is None:
            from .kernels import CudaKernels          # raises when the CUDA library / device is missing: no CPU fallback
            dev = self.device
            if dev is None:
                dev = f"cuda:{torch.cuda.current_device()}" if torch.cuda.is_available() else "cuda:0"
            self._kernels = CudaKernels(dev)
        return self._kernels

    def run(self, initial_state: State, operator, sim_params=None, noise_model: NoiseModel | None = None, *, observables=None,
            num_traj: int | None = None, random_seed: int | None = None, get_state: bool = False) -> Result:
        from .program import SimulationProgram, run_program
        program_kwargs_set = observables is not None or num_traj is not None or random_seed is not None or get_state     # :1222-1260
        if (isinstance(operator, Sequence) and not isinstance(operator, (str, bytes))
                and (not operator or (isinstance(operator[0], tuple) and len(operator[0]) in {2, 3}))):
            if sim_params is not None:
                raise TypeError("sim_params must be None when operator is a segment pair list.")
            operator = SimulationProgram(operator, observables=observables, num_traj=num_traj, random_seed=random_seed, get_state=get_state)
        elif program_kwargs_set and not isinstance(operator, SimulationProgram):
            raise TypeError("observables=, num_traj=, random_seed=, and get_state= are only valid for a SimulationProgram or a segment pair list.")
        if isinstance(operator, SimulationProgram):
            if sim_params is not None:
                raise TypeError("sim_params must be None when operator is a SimulationProgram.")
            if not isinstance(initial_state, State):
                raise TypeError("SimulationProgram execution requires a single State initial_state.")
            return run_program(self, initial_state, operator, noise_model)
        if sim_params is None:
            raise TypeError("Standalone simulation requires AnalogSimParams or DigitalSimParams.")
        if not isinstance(sim_params, (AnalogSimParams, DigitalSimParams)):
            raise TypeError(f"sim_params must be AnalogSimParams or DigitalSimParams, got {type(sim_params).__name__}.")
        if isinstance(initial_state, list):
            raise NotImplementedError("list[State] unitary ensembles are outside the B200 hot-path scope (SURVEY.md 2.1 #3)")
        if not isinstance(initial_state, State):
            raise TypeError("Analog simulation requires initial_state to be a list or State.")
        if noise_model is not None:
            noise_model = noise_model.sample(rng=make_disorder_rng(getattr(sim_params, "random_seed", None)))   # :1269-1271
        result = Result(sim_params=sim_params, noise_model=noise_model)
        if any(o.gate.name == "schmidt_spectrum" for o in (sim_params.observables or [])):
            # the reference's Simulator.run cannot store a vector-valued observable: _store_observable_trajectory assigns the 500-vector
            # into a float row (simulator.py:932-942) and numpy raises exactly this.  The spectrum is available where the reference
            # offers it: backends.analog_tjm_1 (dtype=object results, analog_tjm.py:424-426)
            raise ValueError("setting an array element with a sequence.")
        if isinstance(sim_params, AnalogSimParams):
            if not sim_params.get_state and not sim_params.observables:
                raise ValueError("No output specified: either observables or get_state must be set.")
            if not isinstance(operator, Hamiltonian):
                raise TypeError("Analog simulation requires a Hamiltonian operator.")
            self._run_analog(initial_state, operator, sim_params, noise_model, result)
            return result
        if not sim_params.get_state and not sim_params.observables and sim_params.shots is None:
            raise ValueError("No output specified: set observables, shots, and/or get_state.")        # simulator.py:1283-1290
        from .circuits import Circuit
        from .qasm import load_circuit
        operator = load_circuit(operator)                # OpenQASM 2 text or a path to it (simulator.py:1262-1263, qasm_utils.py:133-162)
        if not isinstance(operator, Circuit) and not (hasattr(operator, "num_qubits") and hasattr(operator, "data")):
            raise TypeError("Circuit simulation requires a QuantumCircuit operator.")                 # :1304-1306 (Circuit or a qiskit circuit)
        from .digital import run_digital                # digital TJM (digital/digital_tjm.py:636-749)
        return run_digital(self, initial_state, operator, sim_params, noise_model, result)

    # ------------------------------------------------------------------------------------------------------------------
    def _run_analog(self, state: State, ham: Hamiltonian, sp: AnalogSimParams, noise: NoiseModel | None, result: Result) -> None:
        run = AnalogRun(self, state, ham, sp, noise, result, resident=False)
        run.advance(None)
        run.result()

    def open(self, initial_state: State, operator: Hamiltonian, sim_params: AnalogSimParams, noise_model: NoiseModel | None = None) -> "AnalogRun":
        """``run`` in incremental form, for callers that stream: the same validation and set-up as ``run`` (state, MPO and
        tables go to the device here), then ``AnalogRun.advance(n)`` executes n time steps -- every step uploads its own uniforms
        from pinned host memory and downloads its result column -- and ``AnalogRun.result()`` aggregates what has been computed.
        ``run(...)`` is ``open(...)`` + ``advance(None)`` + ``result()``, wave by wave when the ensemble does not fit in HBM;
        ``open`` itself keeps this rank's whole shard resident and raises if it cannot."""
        if not isinstance(sim_params, AnalogSimParams):
            raise TypeError("Simulator.open drives the analog TJM (AnalogSimParams)")
        if not isinstance(initial_state, State):
            raise TypeError("Analog simulation requires initial_state to be a list or State.")
        if not isinstance(operator, Hamiltonian):
            raise TypeError("Analog simulation requires a Hamiltonian operator.")
        if not sim_params.get_state and not sim_params.observables:
            raise ValueError("No output specified: either observables or get_state must be set.")
        if noise_model is not None:
            noise_model = noise_model.sample(rng=make_disorder_rng(getattr(sim_params, "random_seed", None)))   # :1269-1271
        return AnalogRun(self, initial_state, operator, sim_params, noise_model, Result(sim_params=sim_params, noise_model=noise_model), resident=True)

    def _finish(self, result: Result, sp, res_local, diag_local, status_local, mine, ntraj: int, nobs: int, ncol: int, to_sorted, times, stats,
                dev) -> None:
        """Combine the shards -- ONE all-reduce of per-column sums + counts (NCCL over NVLink on GPU ranks, gloo on CPU);
        per-trajectory rows are gathered for Result.trajectories -- and fill Result in the user's observable order
        (simulator.py:932-942, result.py:112-138).  Flagged trajectories (non-zero device status): ``on_trajectory_error="raise"``
        raises the exception the reference raises in that situation with the complete Result attached; ``"flag"`` returns the
        Result with ``status`` set and the means taken over the clean trajectories."""
        import torch.distributed as dist
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        good = (status_local == 0) if self.on_trajectory_error == "flag" else np.ones(len(mine), dtype=bool)
        sums = torch.from_numpy(np.concatenate([res_local[good].sum(axis=0).reshape(-1), diag_local[good].sum(axis=0).reshape(-1),
                                                [float(good.sum())]]))
        if world > 1:
            backend = dist.get_backend()
            buf = sums.to(dev) if backend == "nccl" else sums
            dist.all_reduce(buf, op=dist.ReduceOp.SUM)
            sums = buf.cpu()
            gathered: list = [None] * world
            dist.all_gather_object(gathered, (list(mine), res_local, diag_local, status_local))
            res_all = np.zeros((ntraj, nobs, ncol))
            diag_all = np.zeros((ntraj, 3, ncol))
            status_all = np.zeros(ntraj, dtype=np.int32)
            for ids, r_, d_, s_ in gathered:
                if len(ids):
                    res_all[ids], diag_all[ids], status_all[ids] = r_, d_, s_
        else:
            res_all, diag_all, status_all = res_local, diag_local, status_local
        sums = sums.numpy()
        count = max(sums[-1], 1.0)
        nres = nobs * ncol
        mean_res = (sums[:nres] / count).reshape(nobs, ncol)
        mean_diag = (sums[nres: nres + 3 * ncol] / count).reshape(3, ncol)
        result.observables = list(sp.observables)
        result.times = times
        result.trajectories = [res_all[:, to_sorted[i], :].copy() for i in range(len(sp.observables))]
        result.expectation_values = [mean_res[to_sorted[i]].copy() for i in range(len(sp.observables))]
        result.runtime_cost, result.max_bond, result.total_bond = mean_diag[0], mean_diag[1], mean_diag[2]
        result.status = status_all
        result.stats = stats
        if self.on_trajectory_error == "raise":
            exc = status_exception(status_all, result)
            if exc is not None:
                raise exc

    def _wave_size(self, n: int, length: int, ep: EngineParams, mpo, init, dev, bond_factor=None) -> int:
        if n == 0:
            return 0
        if self.batch_size is not None:
            return max(1, min(n, self.batch_size))
        init_bonds = [1] * (length + 1)
        for i, t in enumerate(init):
            init_bonds[i] = max(init_bonds[i], t.shape[1])
            init_bonds[i + 1] = max(init_bonds[i + 1], t.shape[2])
        mpo_bonds = [int(mpo[0].shape[2])] + [int(w.shape[3]) for w in mpo]
        one = BatchLayout(length, 1, layout_chi(ep), mpo_bonds, init_bonds, bond_factor=bond_factor)
        two_envs = ep.tdvp_mode in ("1site", "dynamic")
        try:                                         # the library's own accounting (include/yaqs_b200.h: yb_workspace_bytes)
            from .kernels import KernelError, workspace_bytes
            per = workspace_bytes(length, 1, one.cap, one.D, ep.order, two_envs)["total"]
        except (KernelError, OSError):               # host-twin tests on a box without the built library
            per = bytes_per_trajectory(one, ep.order, two_envs)
        if torch.device(dev).type == "cuda":
            free, _total = torch.cuda.mem_get_info(dev)
        else:
            free = 8 << 30
        return max(1, min(n, int(0.85 * free // per)))


class _LaneRun:
    """One lock-step sub-batch: engine, the generator over its time grid, pinned host staging for the per-step I/O."""

    def __init__(self, sim: "Simulator", K, w0: int, idx: list, length: int, mpo, ep: EngineParams, procs, obs, init, sched, per_interval, seed):
        dev = K.device
        self.w0, self.idx, self.ep, self.K = w0, idx, ep, K
        self.eng = BatchedTJM(K, length, len(idx), mpo, ep, procs, obs, init, scheduled_jumps=sched, extra_mpos=per_interval or ())
        nt, nb = ep.n_times, len(idx)
        self.ncol = nt if ep.sample_timesteps else 1
        cuda = torch.device(dev).type == "cuda"
        self.tu_host = self.su_host = None
        if sim.rng_mode == "numpy":
            tu, su = numpy_tables(seed, idx, nt, ep.order)
            self.tu_host = torch.from_numpy(tu)
            self.su_host = torch.from_numpy(su) if su is not None else None
            if cuda:
                self.tu_host = self.tu_host.pin_memory()
                self.su_host = self.su_host.pin_memory() if self.su_host is not None else None
            self.tu = torch.zeros_like(self.tu_host, device=dev)
            self.su = torch.zeros_like(self.su_host, device=dev) if self.su_host is not None else None
        else:                                              # device generator: the uniforms are drawn inside the jump-test kernel (no tables, no host input)
            self.tu = self.su = None
            self.eng.philox = philox_source(seed, idx, dev)
        nobs = len(obs)
        self.res_host = torch.zeros((self.ncol, nb, nobs), dtype=torch.float64)
        self.diag_host = torch.zeros((self.ncol, nb, 3), dtype=torch.float64)
        if cuda:
            self.res_host, self.diag_host = self.res_host.pin_memory(), self.diag_host.pin_memory()
        self.log = [] if sim.collect_log else None
        self.gen = self.eng.run_iter(self.tu, self.su, self.log, interval_mpos=per_interval)
        self.j_done = -1
        self.h2d_bytes = self.d2h_bytes = 0
        self.step_events: list = []                        # (j, start, end) CUDA events around the device work of time index j
        self.record_events = sim.record_step_events and cuda
        self.cuda = cuda

    def advance(self, upto: int) -> None:
        """Time indices j_done+1 .. upto: upload the step's uniforms (pinned host -> device), run it, download its column."""
        ep = self.ep
        while self.j_done < upto:
            j = self.j_done + 1
            if self.tu_host is not None:
                c0, c1 = 2 * j, min(2 * j + 2, self.tu_host.shape[1])
                self.tu[:, c0:c1].copy_(self.tu_host[:, c0:c1], non_blocking=True)
                self.h2d_bytes += (c1 - c0) * self.tu_host.shape[0] * 8
                if self.su_host is not None:
                    self.su[:, j].copy_(self.su_host[:, j], non_blocking=True)
                    self.h2d_bytes += self.su_host.shape[0] * 2 * 8
            if self.record_events:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            got = next(self.gen)
            assert got == j, (got, j)
            if self.record_events:
                e1.record()
                self.step_events.append((j, e0, e1))
            if ep.sample_timesteps or j == ep.n_times - 1:
                col = j if ep.sample_timesteps else 0
                self.res_host[col].copy_(self.eng.results[:, :, col], non_blocking=True)
                self.diag_host[col].copy_(self.eng.diag[:, :, col], non_blocking=True)
                self.d2h_bytes += self.res_host[col].numel() * 8 + self.diag_host[col].numel() * 8
            self.j_done = j
        if self.cuda:
            torch.cuda.current_stream(self.K.device).synchronize()          # the step's rows are on the host when advance returns


class AnalogRun:
    """An analog TJM run opened for step-wise execution (``Simulator.open``); ``Simulator.run`` drives the same object."""

    def __init__(self, sim: "Simulator", state: State, ham: Hamiltonian, sp: AnalogSimParams, noise: NoiseModel | None, result: Result,
                 resident: bool):
        self.sim, self.sp, self._result = sim, sp, result
        length = state.length
        if ham.length != length:
            raise ValueError(f"State.length={length} does not match Hamiltonian.length={ham.length}.")
        if noise is not None:
            validate_noise_model_for_run(noise, length, sp)
        self.noiseless = noise is None or all(pr["strength"] == 0 for pr in noise.processes)
        if self.noiseless:
            ntraj = 1                                                                     # :1549-1555
        else:
            if sp.get_state:
                raise ValueError("Cannot return state in noisy analog simulation due to stochastics.")
            ntraj = sp.num_traj
        if sp.tdvp_mode == "dynamic" and sp.max_bond_dim is not None:
            import warnings
            warnings.warn('tdvp_mode="dynamic" with max_bond_dim: where a bond reaches the cap the reference\'s right-to-left hand-over '
                          "(integrators.py:450-466) is not a gauge-invariant TDVP step; this engine takes the projector-splitting step "
                          "instead, so results agree with the reference only while bonds stay below the cap (DESIGN.md 4.13)", stacklevel=4)
        self.ntraj = ntraj
        self.to_sorted = sp.observable_sorted_indices
        self.obs = engine_observables(sp.sorted_observables, length)
        self.ep = ep = EngineParams(dt=sp.dt, n_steps=len(sp.times) - 1, order=2 if sp.order != 1 else 1, max_bond_dim=sp.max_bond_dim,
                                    svd_threshold=sp.svd_threshold, trunc_mode=sp.trunc_mode, krylov_tol=sp.krylov_tol, tdvp_mode=sp.tdvp_mode,
                                    tdvp_sweeps=sp.tdvp_sweeps, sample_timesteps=sp.sample_timesteps)
        self.procs = noise.processes if noise is not None else None
        self.sched = noise.scheduled_jumps if noise is not None else None
        from .program import expand_analog_operator
        if ham.is_piecewise and sp.multi_time_observables:
            raise ValueError("Piecewise Hamiltonians do not support multi_time_observables.")
        operator = expand_analog_operator(ham, sp)                  # simulator.py:1517-1534: MPO, or one MPO per interval
        self.per_interval = [m.tensors for m in operator] if isinstance(operator, tuple) else None
        self.mpo = (operator[0] if isinstance(operator, tuple) else operator).tensors
        self.init = state.mps.tensors
        self.length = length
        self.ncol = ep.n_times if ep.sample_timesteps else 1

        import torch.distributed as dist
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        rank = dist.get_rank() if world > 1 else 0
        self.mine = mine = shard_range(ntraj, rank, world)
        self.K = K = sim._get_kernels()
        dev = K.device
        self.res_local = np.zeros((len(mine), len(self.obs), self.ncol))
        self.diag_local = np.zeros((len(mine), 3, self.ncol))
        self.status_local = np.zeros(len(mine), dtype=np.int32)
        self.stats: dict = {}
        self.final_tensors = None
        self.log_all = [] if sim.collect_log else None
        wave = sim._wave_size(len(mine), length, ep, self.mpo, self.init, dev)
        if resident and wave < len(mine):
            raise ValueError(f"Simulator.open keeps the whole shard resident: {len(mine)} trajectories do not fit on the device ({wave} do); "
                             "use Simulator.run (it works through the ensemble in waves) or fewer trajectories per GPU")
        K.launch_count(reset=True)
        from .lanes import Lanes
        nl = sim.lanes if sim.lanes is not None else (2 if (torch.device(dev).type == "cuda" and wave >= 64) else 1)
        self.lanes = Lanes(dev, nl)
        chunk = max(1, -(-max(wave, 1) // self.lanes.n))
        self.jobs = [(w0, list(mine[w0: w0 + chunk])) for w0 in range(0, len(mine), chunk)]
        self.resident = resident
        self.lane_runs: list[_LaneRun] = []
        if resident:
            self.lane_runs = self.lanes.map(self._open_lane, self.jobs)
        self.j_done = -1
        self.h2d_setup_bytes = sum(int(t.nbytes) for t in self.init) + 3 * sum(int(w.nbytes) for w in self.mpo)

    # ------------------------------------------------------------------------------------------------------------------
    def _open_lane(self, job) -> _LaneRun:
        w0, idx = job
        return _LaneRun(self.sim, self.K, w0, idx, self.length, self.mpo, self.ep, self.procs, self.obs, self.init, self.sched, self.per_interval,
                        self.sp.random_seed)

    def _collect(self, lane: _LaneRun) -> None:
        n, w0 = len(lane.idx), lane.w0
        self.res_local[w0: w0 + n] = lane.res_host.numpy().transpose(1, 2, 0)
        self.diag_local[w0: w0 + n] = lane.diag_host.numpy().transpose(1, 2, 0)
        self.status_local[w0: w0 + n] = lane.eng.status.cpu().numpy()
        for k, v in lane.eng.stats.items():
            self.stats[k] = self.stats.get(k, 0) + v
        if self.log_all is not None:
            self.log_all.append((lane.idx, lane.log))
        if self.sp.get_state and self.noiseless:
            eng = lane.eng
            self.final_tensors = (eng.psi if self.ep.order == 2 and self.ep.n_times > 1 else eng.phi).download(0)

    @property
    def n_times(self) -> int:
        return self.ep.n_times

    def advance(self, nsteps: int | None = 1, serial: bool | str = False, between=None) -> int:
        """Execute the next ``nsteps`` time indices (None: all that remain) for every trajectory of this rank; returns the last
        time index done.  Index 0 is the initial measurement, index j the state at ``times[j]``.  ``serial``: lanes one after
        the other instead of concurrently (instrumented steps of bench.py)."""
        last = self.ep.n_times - 1
        target = last if nsteps is None else min(last, self.j_done + int(nsteps))
        if target <= self.j_done:
            return self.j_done
        if self.resident:
            self.lanes.map(lambda lane: lane.advance(target), self.lane_runs, serial=serial, between=between)
        else:                                              # wave by wave: every job runs its whole time grid, then frees its engine
            if target != last:
                raise ValueError("step-wise execution needs Simulator.open (resident shard)")

            def run_job(job):
                lane = self._open_lane(job)
                lane.advance(last)
                self._collect(lane)
                return None

            self.lanes.map(run_job, self.jobs)
        self.j_done = target
        return target

    def io_bytes(self) -> tuple[int, int]:
        """(host -> device, device -> host) bytes moved by the steps executed so far (set-up uploads excluded)."""
        return sum(l.h2d_bytes for l in self.lane_runs), sum(l.d2h_bytes for l in self.lane_runs)

    def result(self) -> Result:
        """Aggregate what has been computed (all ranks must call it: one all-reduce)."""
        sp, sim = self.sp, self.sim
        if self.resident:
            self.stats = {}
            for lane in self.lane_runs:
                self._collect(lane)
        self.stats["kernel_launches"] = self.K.launch_count()
        sim.last_log = self.log_all
        times = np.asarray(sp.times, dtype=np.float64) if sp.sample_timesteps else np.asarray([sp.elapsed_time], dtype=np.float64)
        sim._finish(self._result, sp, self.res_local, self.diag_local, self.status_local, self.mine, self.ntraj, len(self.obs), self.ncol,
                    self.to_sorted, times, self.stats, self.K.device)
        if self.final_tensors is not None:
            self._result.output_state = State(tensors=self.final_tensors)
        return self._result
