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
from .rng import make_disorder_rng, numpy_tables, philox_tables


def shard_range(n: int, rank: int, world: int) -> range:
    """Contiguous trajectory-index shard of rank `rank`."""
    per = math.ceil(n / world) if world > 0 else n
    return range(min(n, rank * per), min(n, (rank + 1) * per))


def validate_noise_model_for_run(noise: NoiseModel, length: int, sim_params=None) -> None:
    """core/data_structures/noise_model.py:668-780 (the MPS-analog checks, scheduled jumps included)."""
    if noise.scheduled_jumps:
        if not isinstance(sim_params, AnalogSimParams):
            raise ValueError("scheduled_jumps are only supported for single-State analog MPS TJM; they are not supported for MCWF, "
                             "Lindblad, digital, or list[State] ensemble runs.")
        if sim_params.order != 1:
            raise ValueError(f"scheduled_jumps are only supported for AnalogSimParams(order=1); got order={sim_params.order}. Order-2 TJM "
                             "applies deterministic jumps inconsistently on the sampling versus trajectory MPS.")
        times = np.asarray(sim_params.times, dtype=float)
        atol = sim_params.dt * 1e-3
        for jump in noise.scheduled_jumps:
            want = 2 ** len(jump["sites"])
            if np.asarray(jump["matrix"]).shape != (want, want):
                raise ValueError(f"Scheduled jump '{jump['name']}': matrix must be {want}x{want} for qubit chains.")
            for s_ in jump["sites"]:
                if s_ >= length:
                    raise ValueError(f"Scheduled jump '{jump['name']}' acts on site {s_} but the chain has {length} sites.")
            if not np.any(np.isclose(times, float(jump["time"]), atol=atol, rtol=0.0)):
                raise ValueError(f"Scheduled jump time {float(jump['time'])} is not on the simulation time grid (atol={atol}, rtol=0).")
    for pr in noise.processes:
        for s in pr["sites"]:
            if s >= length:
                raise ValueError(f"Noise process '{pr['name']}' acts on site {s} but the chain has {length} sites.")
        if len(pr["sites"]) == 1 and np.asarray(pr["matrix"]).shape != (2, 2):
            raise ValueError(f"Noise process '{pr['name']}': one-site matrix must be 2x2 for qubit chains.")
        if len(pr["sites"]) == 2 and "matrix" in pr and np.asarray(pr["matrix"]).shape != (4, 4):
            raise ValueError(f"Noise process '{pr['name']}': two-site matrix must be 4x4 for qubit chains.")


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
                 batch_size: int | None = None, rng_mode: str = "numpy", collect_log: bool = False, lanes: int | None = None) -> None:
        self.parallel, self.max_workers, self.show_progress = parallel, max_workers, show_progress
        self.device = device
        self._kernels = kernels
        self.batch_size = batch_size
        if rng_mode not in ("numpy", "philox"):
            raise ValueError("rng_mode must be 'numpy' (reference-identical seeded streams) or 'philox' (device generator)")
        self.rng_mode = rng_mode
        self.collect_log = collect_log
        self.last_log: list | None = None
        self.lanes = lanes          # concurrent sub-batches per GPU (None: 2 on CUDA for waves of >= 64 trajectories, else 1)

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
        if isinstance(sim_params, AnalogSimParams):
            if not sim_params.get_state and not sim_params.observables:
                raise ValueError("No output specified: either observables or get_state must be set.")
            if not isinstance(operator, Hamiltonian):
                raise TypeError("Analog simulation requires a Hamiltonian operator.")
            self._run_analog(initial_state, operator, sim_params, noise_model, result)
            return result
        from .digital import run_digital                # digital TJM (digital/digital_tjm.py:636-749)
        return run_digital(self, initial_state, operator, sim_params, noise_model, result)

    # ------------------------------------------------------------------------------------------------------------------
    def _run_analog(self, state: State, ham: Hamiltonian, sp: AnalogSimParams, noise: NoiseModel | None, result: Result) -> None:
        length = state.length
        if ham.length != length:
            raise ValueError("State and Hamiltonian must have the same number of sites.")
        if noise is not None:
            validate_noise_model_for_run(noise, length, sp)
        noiseless = noise is None or all(pr["strength"] == 0 for pr in noise.processes)
        if noiseless:
            ntraj = 1                                                                     # :1549-1555
        else:
            if sp.get_state:
                raise ValueError("Cannot return state in noisy analog simulation due to stochastics.")
            ntraj = sp.num_traj
        if sp.tdvp_mode == "dynamic" and sp.max_bond_dim is not None:
            import warnings
            warnings.warn('tdvp_mode="dynamic" with max_bond_dim: where a bond reaches the cap the reference\'s right-to-left hand-over '
                          "(integrators.py:450-466) is not a gauge-invariant TDVP step; this engine takes the projector-splitting step "
                          "instead, so results agree with the reference only while bonds stay below the cap (DESIGN.md 4.13)", stacklevel=3)
        sorted_obs = sp.sorted_observables
        to_sorted = sp.observable_sorted_indices
        obs = engine_observables(sorted_obs, length)
        ep = EngineParams(dt=sp.dt, n_steps=len(sp.times) - 1, order=2 if sp.order != 1 else 1, max_bond_dim=sp.max_bond_dim,
                          svd_threshold=sp.svd_threshold, trunc_mode=sp.trunc_mode, krylov_tol=sp.krylov_tol, tdvp_mode=sp.tdvp_mode,
                          tdvp_sweeps=sp.tdvp_sweeps, sample_timesteps=sp.sample_timesteps)
        procs = noise.processes if noise is not None else None
        sched = noise.scheduled_jumps if noise is not None else None
        from .program import expand_analog_operator
        if ham.is_piecewise and sp.multi_time_observables:
            raise ValueError("Piecewise Hamiltonians do not support multi_time_observables.")
        operator = expand_analog_operator(ham, sp)                  # simulator.py:1517-1534: MPO, or one MPO per interval
        per_interval = [m.tensors for m in operator] if isinstance(operator, tuple) else None
        mpo = (operator[0] if isinstance(operator, tuple) else operator).tensors
        init = state.mps.tensors
        ncol = ep.n_times if ep.sample_timesteps else 1

        import torch.distributed as dist
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        rank = dist.get_rank() if world > 1 else 0
        mine = shard_range(ntraj, rank, world)

        K = self._get_kernels()
        dev = K.device
        res_local = np.zeros((len(mine), len(obs), ncol))
        diag_local = np.zeros((len(mine), 3, ncol))
        status_local = np.zeros(len(mine), dtype=np.int32)
        stats: dict = {}
        final_tensors = None
        log_all = [] if self.collect_log else None
        wave = self._wave_size(len(mine), length, ep, mpo, init, dev)
        K.launch_count(reset=True)
        from .lanes import Lanes
        nl = self.lanes if self.lanes is not None else (2 if (torch.device(dev).type == "cuda" and wave >= 64) else 1)
        lanes = Lanes(dev, nl)
        chunk = max(1, -(-max(wave, 1) // lanes.n))
        jobs = [(w0, list(mine[w0: w0 + chunk])) for w0 in range(0, len(mine), chunk)]

        def run_job(job):
            w0, idx = job
            eng = BatchedTJM(K, length, len(idx), mpo, ep, procs, obs, init, scheduled_jumps=sched, extra_mpos=per_interval or ())
            if self.rng_mode == "numpy":
                tu, su = numpy_tables(sp.random_seed, idx, ep.n_times, ep.order)
                tu = torch.from_numpy(tu).to(dev)
                su = torch.from_numpy(su).to(dev) if su is not None else None
            else:
                tu, su = philox_tables(K, sp.random_seed, idx, ep.n_times, ep.order, dev)
            log = [] if self.collect_log else None
            r, dg = eng.run(tu, su, log, interval_mpos=per_interval)
            fin = None
            if sp.get_state and noiseless:
                fin = (eng.psi if ep.order == 2 and ep.n_times > 1 else eng.phi).download(0)
            return w0, idx, r.cpu().numpy(), dg.cpu().numpy(), eng.status.cpu().numpy(), dict(eng.stats), log, fin

        for w0, idx, r, dg, stt, est, log, fin in lanes.map(run_job, jobs):
            res_local[w0: w0 + len(idx)] = r
            diag_local[w0: w0 + len(idx)] = dg
            status_local[w0: w0 + len(idx)] = stt
            for k, v in est.items():
                stats[k] = stats.get(k, 0) + v
            if log_all is not None:
                log_all.append((idx, log))
            if fin is not None:
                final_tensors = fin
        stats["kernel_launches"] = K.launch_count()
        self.last_log = log_all

        times = np.asarray(sp.times, dtype=np.float64) if sp.sample_timesteps else np.asarray([sp.elapsed_time], dtype=np.float64)
        self._finish(result, sp, res_local, diag_local, status_local, mine, ntraj, len(obs), ncol, to_sorted, times, stats, dev)
        if final_tensors is not None:
            result.output_state = State(tensors=final_tensors)

    def _finish(self, result: Result, sp, res_local, diag_local, status_local, mine, ntraj: int, nobs: int, ncol: int, to_sorted, times, stats,
                dev) -> None:
        """Combine the shards -- ONE all-reduce of per-column sums + counts (NCCL over NVLink on GPU ranks, gloo on CPU);
        per-trajectory rows are gathered for Result.trajectories -- and fill Result in the user's observable order
        (simulator.py:932-942, result.py:112-138)."""
        import torch.distributed as dist
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        sums = torch.from_numpy(np.concatenate([res_local.sum(axis=0).reshape(-1), diag_local.sum(axis=0).reshape(-1), [float(len(mine))]]))
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
        count = sums[-1]
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
        if status_all.any():
            bad = np.nonzero(status_all)[0]
            raise RuntimeError(f"device status words non-zero for trajectories {bad[:8].tolist()} (bit 1: Jacobi SVD not converged, "
                               "bit 2: non-finite singular values, bit 4: zero jump weights)")

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
        per = bytes_per_trajectory(one, ep.order, ep.tdvp_mode in ("1site", "dynamic"))
        if torch.device(dev).type == "cuda":
            free, _total = torch.cuda.mem_get_info(dev)
        else:
            free = 8 << 30
        return max(1, min(n, int(0.85 * free // per)))
