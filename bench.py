#!/usr/bin/env python
"""bench.py -- noisy-MPS TJM throughput on B200 (BASELINE.json metric) with roofline and CPU baseline.

    python bench.py --gpus N --steps K --warmup W              # our arm (torchrun launches one rank per GPU for N > 1)
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU arm: the oracle port on the host cores
    python bench.py --workload C1|C2|C3|C4|C5 ...                  # other BASELINE.json configs (C3 is the headline and the default)

Workload C3 = BASELINE.json configs[2]: transverse-field Ising chain L=100, chi_max=128, relaxation+dephasing (lowering +
pauli_z, gamma=0.1 on every site), second-order TJM with Z on every site sampled every step, started from a seeded random MPS
whose bonds are saturated at chi (otherwise the first ~50 steps run at tiny chi and measure nothing).

ONE pass through the public API: ``Simulator.open(State, Hamiltonian, AnalogSimParams, NoiseModel)`` uploads the state, then
every step is one ``AnalogRun.advance(1)`` = upload of that step's uniforms from pinned host memory, step_through (TDVP sweep,
dissipation, jump test) + sample (copy, TDVP sweep, dissipation, jump test, diagnostics, observables) for every trajectory of the
rank, download of that step's result column into pinned host memory.  W warm-up steps, then exactly K timed steps between
barriers.  From the SAME K steps:
  value  trajectory-steps/s on the device: CUDA events around the device work of every step (copies and host gaps excluded)
  e2e    trajectory-steps/s by the wall clock around the K advance() calls (host buffers in, host buffers out)
Weak scaling: the batch per GPU is fixed; value = (batch * N * K) / max-over-ranks time.  The run ends with the one all-reduce
of the path (per-column observable sums + counts) inside ``AnalogRun.result()``; status words and finiteness are checked on it.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNIT = "trajectory-steps/s"
FP64_PEAK_TFLOPS = 37.1      # measured on this pool's B200: profiles/r01_fp64_peak.json (DMMA == DFMA issue peak)

# BASELINE.json configs as bench workloads.  batch = trajectories per GPU, sized so that the driver's
# `--steps 20 --warmup 5` finishes within its time box; "saturated" = seeded random MPS with every bond at min(chi, 2^k).
WORKLOADS = {
    "C1": dict(model="ising", length=10, chi=16, noise=("lowering", "pauli_z"), gamma=0.1, batch=4096, threshold=1e-6, bound="hbm",
               text="Ising L=10, chi=16"),
    "C2": dict(model="heisenberg", length=30, chi=64, noise=("lowering",), gamma=0.1, batch=1000, threshold=1e-6, bound="tensor", zz=True,
               text="Heisenberg L=30, chi=64"),
    "C3": dict(model="ising", length=100, chi=128, noise=("lowering", "pauli_z"), gamma=0.1, batch=192, threshold=1e-6, bound="tensor",
               text="Ising L=100, chi=128"),
    # BASELINE.json configs[3]: digital noisy Trotterized Ising circuit, 49 qubits, 2-qubit depolarizing noise as the 15 two-qubit Paulis
    # per pair (9 crosstalk_ab per pair + pauli_x/y/z per site, tests/golden dig_ising49_c4_full); a step = one Trotter step of the circuit
    "C4": dict(kind="digital", model="ising", length=49, chi=64, gamma=1e-3, batch=2048, threshold=1e-6, bound="hbm", text="digital Ising circuit, 49 qubits"),
    "C5": dict(model="heisenberg", length=200, chi=256, noise=("lowering", "pauli_z"), gamma=0.05, xx=0.01, batch=16, threshold=1e-6, bound="tensor",
               text="Heisenberg L=200, chi=256"),
}


def metric_name(wl: dict) -> str:
    if wl.get("kind") == "digital":
        return f"noisy-MPS trajectory-steps/s (digital Trotterized Ising circuit, {wl['length']} qubits, chi_max={wl['chi']}, digital TJM, one step = one Trotter step)"
    name = "Ising" if wl["model"] == "ising" else "Heisenberg"
    return f"noisy-MPS trajectory-steps/s ({name} L={wl['length']}, chi={wl['chi']}, order-2 TJM, sampled every step)"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("YB_BENCH_WORKLOAD", "C3"), choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=int(os.environ.get("YB_BENCH_BATCH", "0")), help="trajectories per GPU (0: the workload's default)")
    ap.add_argument("--lanes", type=int, default=int(os.environ.get("YB_BENCH_LANES", "0")),
                    help="concurrent lock-step sub-batches per GPU (host thread + stream each); 0 = 3 when the box has at least 4 host cores per "
                         "rank, else 2 (measured at C3: 3 x 64 trajectories 14.1 s per step, 2 x 96 14.5 s; every lane is a busy host thread)")
    ap.add_argument("--length", type=int, default=0)
    ap.add_argument("--chi", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--roofline-steps", type=int, default=1, help="in the last n timed steps ONE lane runs alone with CUDA events around "
                    "every launch (0: none), the other lanes follow concurrently.  Such a step is slower than a fully concurrent one and it "
                    "counts in `value` and `e2e` like every other timed step")
    ap.add_argument("--cpu-cores", type=int, default=0, help="cores for the CPU arm (0 = all)")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.length:
        wl["length"] = args.length
    if args.chi:
        wl["chi"] = args.chi
    if args.batch:
        wl["batch"] = args.batch
    wl["name"] = args.workload
    args.wl = wl
    if args.lanes <= 0:
        world = max(1, int(os.environ.get("WORLD_SIZE", "1")))
        args.lanes = 3 if (os.cpu_count() or 1) // world >= 4 else 2
    return args


# ----------------------------------------------------------------------------------------------------------------------
# synthetic workload (host side)
# ----------------------------------------------------------------------------------------------------------------------
def saturated_tensors(length: int, chi: int, seed: int = 1234):
    """Seeded random MPS with bonds min(chi, 2**k); normalised afterwards (MPS.normalize / the oracle's normalize)."""
    rng = np.random.default_rng(seed)
    bonds = [1] + [min(chi, 2 ** min(i + 1, length - 1 - i)) for i in range(length - 1)] + [1]
    # every tensor is scaled to a transfer-matrix norm of order one: the state is normalised afterwards anyway, but the raw norm of a
    # product of 200 unscaled 2 x 256 x 256 Gaussian tensors (~1e271, squared in the norm) overflows a double (config 5)
    return [((rng.standard_normal((2, bonds[i], bonds[i + 1])) + 1j * rng.standard_normal((2, bonds[i], bonds[i + 1]))) / np.sqrt(4.0 * bonds[i]))
            .astype(np.complex128) for i in range(length)]


def noise_processes(wl: dict) -> list[dict]:
    L = wl["length"]
    procs = [{"name": n, "sites": [i], "strength": wl["gamma"]} for i in range(L) for n in wl["noise"]]
    if wl.get("xx"):
        procs += [{"name": "crosstalk_xx", "sites": [i, i + 1], "strength": wl["xx"]} for i in range(L - 1)]
    return procs


def workload_config(args, nranks):
    wl = args.wl
    return {"workload": wl["name"], "hamiltonian": "ising J=1 g=0.5" if wl["model"] == "ising" else "heisenberg Jx=Jy=Jz=1",
            "length": wl["length"], "chi_max": wl["chi"], "noise": "+".join(wl["noise"]) + f" gamma={wl['gamma']} per site" + (f", crosstalk_xx {wl['xx']} per pair" if wl.get("xx") else ""),
            "order": 2, "dt": 0.1, "svd_threshold": wl["threshold"], "krylov_tol": 1e-4,
            "observables": "Z on every site" + (" + ZZ on every pair" if wl.get("zz") else "") + ", every step",
            "initial_state": "seeded random MPS, bonds saturated at chi",
            "batch_per_gpu": wl["batch"], "global_batch": wl["batch"] * nranks, "sharding": f"trajectories x{nranks}", "lanes_per_gpu": args.lanes,
            "l2_policy": "inputs larger than L2 (per-step working set >> 126 MB)" if wl["chi"] >= 64 else "state + environments of the batch exceed L2",
            "library_options": {k: os.environ[k] for k in ("YB_GEMM_RAGGED", "YB_WAPPLY_ONEPASS", "YB_HEEV_RING", "YB_HEEV_QL", "YB_HETRD_BLOCKED", "YB_GEMM_SMALL_MAX", "YB_CHOLQR", "YB_HEEV_TOP", "YB_GRAPHS") if k in os.environ}}


# ----------------------------------------------------------------------------------------------------------------------
# SURVEY.md 8(d) flop model with real counters
# ----------------------------------------------------------------------------------------------------------------------
def flop_model_step(bonds, mpo_bonds, n_mv2: float, n_mv1: float, n_jumps: float, d: int = 2) -> dict:
    """Algorithmic flops of ONE order-2 sampled TJM step of ONE trajectory by the normative model of SURVEY.md 8(d)
    (8 real flops per complex multiply-add), evaluated on the measured mean bond profile `bonds` (length L + 1), the MPO
    bonds `mpo_bonds` (length L + 1) and the measured matvec / jump counts per trajectory-step.  Per-site terms are averaged
    over the chain and multiplied by the call counts of integrators.py:161-291 / dissipation.py / stochastic_process.py."""
    L = len(bonds) - 1

    def mv(dt, i, ns):                                   # H_eff matvec on sites i .. i+ns-1 (primitives.py:180-204)
        cl, cr, Dl, Dr = bonds[i], bonds[i + ns], mpo_bonds[i], mpo_bonds[i + ns]
        return 8.0 * (dt * cl * cr * (Dr * cr + Dl * cl) + dt * dt * Dl * Dr * cl * cr)

    def svd(m, n):                                       # complex economy SVD, Golub-Van Loan count x 4
        m, n = max(m, n), min(m, n)
        return 4.0 * (14.0 * m * n * n + 8.0 * n ** 3)

    def qr(m, n):
        m, n = max(m, n), min(m, n)
        return 8.0 * (2.0 * m * n * n - 2.0 / 3.0 * n ** 3)

    mean = lambda xs: float(sum(xs)) / max(len(xs), 1)      # noqa: E731
    mv2 = mean([mv(d * d, i, 2) for i in range(L - 1)])
    mv1 = mean([mv(d, i, 1) for i in range(L)])
    merge = mean([8.0 * d * d * bonds[i] * bonds[i + 1] * bonds[i + 2] for i in range(L - 1)])
    split = mean([svd(d * bonds[i], d * bonds[i + 2]) for i in range(L - 1)])
    shift = mean([svd(d * bonds[i], bonds[i + 1]) + 8.0 * d * bonds[i] * bonds[i + 1] ** 2 for i in range(1, L)])
    jump = L * mean([qr(d * bonds[i], bonds[i + 1]) for i in range(L)]) + (L - 1) * shift
    sweep_krylov = n_mv2 * mv2 + n_mv1 * mv1                 # both sweeps of the step (phi and psi) are in the counters
    env = 2 * (3 * L - 4) * mv1
    splits = 2 * (2 * L - 3) * (merge + split)
    shifts = 4 * (L - 1) * shift
    total = sweep_krylov + env + splits + shifts + n_jumps * jump
    return {"total": total, "krylov": sweep_krylov, "environments": env, "merge_and_split": splits, "gauge_shifts": shifts, "jumps": n_jumps * jump}


def hbm_bytes_step(bonds, mpo_bonds, d: int = 2) -> float:
    """SURVEY.md 8(d), HBM-bound regime: minimum traffic of one order-2 sampled step of one trajectory = 15 S + 4 E with
    S = sum_sites 16 d chi_l chi_r (state bytes) and E = sum_bonds 16 chi^2 D (environment bytes)."""
    L = len(bonds) - 1
    S = sum(16.0 * d * bonds[i] * bonds[i + 1] for i in range(L))
    E = sum(16.0 * bonds[j] * bonds[j] * mpo_bonds[j] for j in range(L + 1))
    return 15.0 * S + 4.0 * E


# ----------------------------------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index: int):
        self.index, self.samples, self.reasons, self._stop = index, [], set(), threading.Event()
        self.max_mhz = None
        self.thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if "Active" in v and "Not" not in v:
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop.wait(0.5)

    def __enter__(self):
        self.thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self.thread.join(timeout=5)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port on the host cores (reference = pure Python + numpy/scipy; the oracle restates it)
# ----------------------------------------------------------------------------------------------------------------------
CPU_SAMPLE_LENGTHS = (24, 40)     # two chains of the workload's chi; the per-site slope between them extrapolates to the full chain


def _cpu_half_step(job):
    """One step_through (TDVP sweep + dissipation(dt) + jump test) of ONE trajectory = half a sampled order-2 step."""
    wl, length, traj = job
    from threadpoolctl import threadpool_limits

    from oracle import tjm_oracle as O
    with threadpool_limits(1):
        chi = wl["chi"]
        state = O.OracleMPS(saturated_tensors(length, chi), None)
        O.normalize(state, "B", "QR")
        p = O.OracleParams(dt=0.1, n_steps=1, order=2, max_bond_dim=chi, svd_threshold=wl["threshold"], krylov_tol=1e-4)
        mpo = O.ising_mpo(length, 1.0, 0.5) if wl["model"] == "ising" else O.heisenberg_mpo(length, 1.0, 1.0, 1.0)
        noise = [O.make_process(n, [i], wl["gamma"]) for i in range(length) for n in wl["noise"]]
        if wl.get("xx"):
            noise += [O.make_process("crosstalk_xx", [i, i + 1], wl["xx"]) for i in range(length - 1)]
        rng = O.UniformStream(O.trajectory_stream(42, traj))
        t0 = time.perf_counter()
        O.tdvp(state, mpo, p)
        O.apply_dissipation(state, noise, p.dt, p)
        O.stochastic_process(state, noise, p.dt, p, rng)
        return time.perf_counter() - t0


def cpu_pool_wall(wl: dict, length: int, cores: int, first_traj: int = 0) -> float:
    """All `cores` workers run one half-step of a chain of `length` sites each, concurrently (the reference's executor is a fork
    pool with one BLAS thread per worker, core/parallel_utils.py:307-358).  Returns the wall seconds of the pool."""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_half_step, [(wl, length, first_traj + t) for t in range(cores)])
    return time.perf_counter() - t0


def cpu_extrapolate(wl: dict, walls: dict, cores: int) -> tuple[float, float]:
    """(trajectory-steps/s on `cores` cores for the full chain, wall seconds of one full-chain half-step) from the pool walls
    of the two sample chains: bulk sites of a saturated chain all cost the same, so the half-step time is linear in the chain
    length beyond the bond ramps -- t(L) = t(b) + (L - b) * (t(b) - t(a)) / (b - a).  Checked here against full L = 100 chains
    of C3 (profiles/r02_cpu_sample_extrapolation.json): 28.8 s predicted, 29.5 s (no jump) / 36.1 s (one jump) measured, i.e. the
    extrapolation favours the CPU."""
    a, b = CPU_SAMPLE_LENGTHS
    L = wl["length"]
    if L <= b:
        t_full = walls[L]
    else:
        t_full = walls[b] + (L - b) * (walls[b] - walls[a]) / (b - a)
    return 0.5 * cores / t_full, t_full


def cpu_sample_text(wl: dict, cores: int) -> str:
    a, b = CPU_SAMPLE_LENGTHS
    if wl["length"] <= b:
        return (f"{cores} trajectories in parallel (fork pool, 1 BLAS thread each), each one step_through (TDVP sweep + dissipation(dt) + jump test) "
                f"of the full chain = 0.5 sampled order-2 step")
    return (f"{cores} trajectories in parallel (fork pool, 1 BLAS thread each), each one step_through (TDVP sweep + dissipation(dt) + jump test) on "
            f"saturated chi={wl['chi']} chains of {a} and {b} sites; extrapolated linearly in the chain length to L={wl['length']} "
            f"(bulk sites cost the same) = 0.5 sampled order-2 step; port of the reference, extrapolated")


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    wl = args.wl
    cores = args.cpu_cores or os.cpu_count() or 1
    small = dict(wl, chi=16)
    for _ in range(args.warmup):
        cpu_pool_wall(small, 10, min(cores, 2))                 # warm the pool / BLAS on a tiny chain (untimed)
    a, b = CPU_SAMPLE_LENGTHS
    lengths = [wl["length"]] if wl["length"] <= b else [a, b]
    t0 = time.perf_counter()
    walls: dict = {n: [] for n in lengths}
    for k in range(args.steps):                                 # step k samples one of the chain lengths, alternating
        n = lengths[k % len(lengths)]
        walls[n].append(cpu_pool_wall(wl, n, cores, first_traj=k * cores))
    if any(not v for v in walls.values()):                      # fewer steps than lengths: take the missing sample once
        for n in lengths:
            if not walls[n]:
                walls[n].append(cpu_pool_wall(wl, n, cores))
    total = time.perf_counter() - t0
    mean_walls = {n: float(np.mean(v)) for n, v in walls.items()}
    value, t_full = cpu_extrapolate(wl, mean_walls, cores)
    line = {"impl": "reference", "metric": metric_name(wl), "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 * total / max(args.steps, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "complex128 (f64)", "data": "synthetic", "config": workload_config(args, 1),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": cpu_sample_text(wl, cores),
                             "pool_wall_s": mean_walls, "full_chain_half_step_s": t_full},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)
    return 0


# ----------------------------------------------------------------------------------------------------------------------
# digital workload (BASELINE.json configs[3])
# ----------------------------------------------------------------------------------------------------------------------
def digital_noise(wl: dict) -> list[dict]:
    """2-qubit depolarizing noise of strength p = gamma per pair as the 15 two-qubit Paulis at p / 15 each: 9 crosstalk_ab per pair
    plus the 6 single-qubit Paulis of the pair, i.e. pauli_x / y / z at 2 p / 15 per site (noise_model.py:52-60; fixture dig_ising49_c4_full)."""
    L, p = wl["length"], wl["gamma"]
    procs = [{"name": f"pauli_{a}", "sites": [i], "strength": 2.0 * p / 15.0} for i in range(L) for a in "xyz"]
    procs += [{"name": f"crosstalk_{a}{b}", "sites": [i, i + 1], "strength": p / 15.0} for i in range(L - 1) for a in "xyz" for b in "xyz"]
    return procs


def digital_config(args, nranks, nsteps):
    wl = args.wl
    return {"workload": wl["name"], "circuit": f"create_ising_circuit(L={wl['length']}, J=1, g=0.5, dt=0.1, timesteps={nsteps}), sampled after every Trotter step",
            "length": wl["length"], "chi_max": wl["chi"], "noise": f"2-qubit depolarizing p={wl['gamma']} per pair (15 two-qubit Paulis at p/15)", "gate_mode": "mpo",
            "svd_threshold": wl["threshold"], "observables": "Z on every qubit, every Trotter step", "initial_state": "|0...0>",
            "batch_per_gpu": wl["batch"], "global_batch": wl["batch"] * nranks, "sharding": f"trajectories x{nranks}", "lanes_per_gpu": 1,
            "l2_policy": "state of the batch exceeds L2 from the first Trotter steps on (2048 trajectories x 49 sites)"}


def _cpu_digital(job):
    """One trajectory of the oracle's digital_tjm over `nsteps` Trotter steps (the reference's per-trajectory worker)."""
    wl, nsteps, traj = job
    from threadpoolctl import threadpool_limits

    from oracle import tjm_oracle as O
    with threadpool_limits(1):
        L = wl["length"]
        layers = O.ising_trotter_layers(L, 1.0, 0.5, 0.1, nsteps, False, sample_each_step=True)
        procs = [O.make_process(pr["name"], pr["sites"], pr["strength"]) for pr in digital_noise(wl)]
        obs = [(O.OBSERVABLE_MATRICES["z"], [s]) for s in range(L)]
        p = O.OracleParams(dt=1.0, n_steps=1, order=1, max_bond_dim=wl["chi"], svd_threshold=wl["threshold"], krylov_tol=1e-4, digital=True)
        rng = O.UniformStream(O.trajectory_stream(42, traj))
        t0 = time.perf_counter()
        O.digital_tjm(O.product_state(L, "zeros"), layers, procs, p, obs, rng, sample_layers=True, gate_mode="mpo")
        return time.perf_counter() - t0


def cpu_digital_value(wl: dict, nsteps: int, cores: int, first_traj: int = 0) -> tuple[float, float]:
    import multiprocessing as mp
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(cores) as pool:
        pool.map(_cpu_digital, [(wl, nsteps, first_traj + t) for t in range(cores)])
    wall = time.perf_counter() - t0
    return cores * nsteps / wall, wall


def run_reference_digital(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    wl = args.wl
    cores = args.cpu_cores or os.cpu_count() or 1
    nsteps = max(args.warmup, 0) + max(args.steps, 1)
    value, wall = cpu_digital_value(wl, nsteps, cores)
    emit({"impl": "reference", "metric": metric_name(wl), "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
          "ms_per_step": 1000.0 * wall / nsteps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "complex128 (f64)",
          "data": "synthetic", "config": digital_config(args, 1, nsteps),
          "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                           "sample": f"{cores} trajectories in parallel (fork pool, 1 BLAS thread each), each the first {nsteps} Trotter steps of the circuit"},
          "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    return 0


def run_b200_digital(args):
    """ONE call of the public API -- Simulator.run(State, circuit, DigitalSimParams(sample_layers=True), NoiseModel) on a circuit of W + K
    Trotter steps; the engine's sample hook marks the end of every Trotter step (CUDA event + a device -> host read of that step's
    observable column), `value` / `e2e` are the K steps after the first W by the device / wall clock."""
    import torch
    import torch.distributed as dist

    from yaqs_b200 import DigitalSimParams, NoiseModel, Observable, Simulator, State, Z
    from yaqs_b200.circuits import create_ising_circuit
    from yaqs_b200.kernels import CudaKernels

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    wl = args.wl
    L, B, K, W = wl["length"], wl["batch"], max(args.steps, 1), max(args.warmup, 0)
    nsteps = W + K
    line = {"metric": metric_name(wl), "value": None, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": None, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "complex128 (f64)", "data": "synthetic", "config": digital_config(args, world, nsteps)}
    t_start = time.perf_counter()
    try:
        kern = CudaKernels(dev)
        circ = create_ising_circuit(L, 1.0, 0.5, 0.1, nsteps, sample_every_step=True)
        sp = DigitalSimParams(observables=[Observable(Z(), s) for s in range(L)], num_traj=B * world, max_bond_dim=wl["chi"], svd_threshold=wl["threshold"],
                              sample_layers=True, random_seed=42, gate_mode="mpo")
        sim = Simulator(show_progress=False, device=dev, kernels=kern, on_trajectory_error="flag")
        marks: list = []
        engines: list = []

        def on_sample(eng, col, results, diag):
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            host = results[:, :, col].cpu()                   # the step's result column, device -> host (synchronises)
            marks.append((col, ev, time.perf_counter(), host.numel() * 8, kern.launch_count()))
            if not engines or engines[-1] is not eng:
                engines.append(eng)
        sim.on_sample = on_sample
        with ClockSampler(local) as clocks:
            res = sim.run(State(L, initial="zeros"), circ, sp, NoiseModel(digital_noise(wl)))
            torch.cuda.synchronize(dev)
        if len(engines) != 1:
            raise RuntimeError(f"the batch was split into {len(engines)} waves; choose --batch so that one wave holds it")
        by_col = {m[0]: m for m in marks}
        a, b = by_col[W], by_col[W + K]                      # column j is recorded at the end of Trotter step j (column 0: the initial state)
        ms_dev, t_wall = a[1].elapsed_time(b[1]), b[2] - a[2]
        t = torch.tensor([ms_dev, t_wall], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_dev, t_wall = float(t[0].item()), float(t[1].item())
        value = B * world * K / (ms_dev / 1000.0)
        line.update(value=value, ms_per_step=ms_dev / K, gpu_launches=int(b[4] - a[4]), clocks=clocks.summary())
        line["e2e"] = {"value": B * world * K / t_wall, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": int(b[3]), "seconds": t_wall,
                       "api": "Simulator.run(State, circuit, DigitalSimParams(sample_layers=True), NoiseModel): one call; the uniforms of the whole circuit "
                              "go host -> device at set-up (2 doubles per two-qubit gate and trajectory), every Trotter step's observable column comes "
                              "back device -> host; same K steps as `value` by the wall clock",
                       "setup_h2d_bytes": int(B * 16 * (L - 1) * nsteps)}
        eng = engines[0]
        prof = eng.phi.bond.to(torch.float64).mean(dim=0).cpu().tolist()
        S = sum(16.0 * 2 * prof[i] * prof[i + 1] for i in range(L))
        by = 4.0 * S * (L - 1)                               # per Trotter step: every two-qubit gate walks the chain twice (dissipation.py:79-183), read + write
        peak, src = 6468.3, "fallback = the figure MEASURED_PEAKS.json held in round 1"
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
                peak, src = float(json.load(fh)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
        except Exception:
            pass
        achieved = by * B * K / (ms_dev / 1000.0) / 1e9
        line["roofline"] = {"kernel": "whole Trotter step (state streaming: L - 1 two-qubit gates x 2 chain walks x read + write of the state = 4 (L - 1) S bytes)",
                            "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None, "peak_source": src,
                            "algorithmic_bytes_per_trajectory_step": by, "mean_bond_final": float(np.mean(prof[1:L])),
                            "note": "bonds stay <= 11 in this circuit: the step is launch / latency-bound, the fraction is reported as measured"}
        line["counters"] = {k: int(v) for k, v in eng.stats.items()}
        zlast = np.array([res.expectation_values[s][-1] for s in range(L)])
        line["consistency"] = {"status_nonzero": int(np.count_nonzero(res.status)), "status_values": sorted({int(v) for v in res.status}),
                               "observables_finite": bool(all(np.isfinite(ev).all() for ev in res.expectation_values)),
                               "max_abs_z": float(np.max(np.abs(zlast))), "z_mean_first_sites": [float(x) for x in zlast[:4]],
                               "trajectories": int(len(res.status)), "columns": int(len(marks))}
        del eng, engines
        torch.cuda.empty_cache()
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            try:
                cores = args.cpu_cores or os.cpu_count() or 1
                v, wall = cpu_digital_value(wl, nsteps, cores)
                line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "pool_wall_s": wall,
                                        "sample": f"{cores} trajectories in parallel (fork pool, 1 BLAS thread each), each the first {nsteps} Trotter steps "
                                                  f"of the circuit (the warm-up steps, whose bonds are smaller, are counted: favours the CPU)"}
            except Exception as exc:  # noqa: BLE001
                line["cpu_baseline"] = {"error": repr(exc)}
    except Exception as exc:  # noqa: BLE001 -- never lose the line
        import traceback
        line["error"] = repr(exc)
        line["traceback"] = traceback.format_exc()[-1500:]
    line["bench_wall_s"] = time.perf_counter() - t_start
    if rank == 0:
        emit(line)
    if world > 1:
        try:
            dist.destroy_process_group()
        except Exception:
            pass
    return 0 if "error" not in line else 1


# ----------------------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------------------
def run_b200_arm(args):
    import torch
    import torch.distributed as dist

    from yaqs_b200 import AnalogSimParams, Hamiltonian, NoiseModel, Observable, Simulator, State, Z, ZZ
    from yaqs_b200.kernels import CudaKernels

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    wl = args.wl
    L, chi, B, K, W = wl["length"], wl["chi"], wl["batch"], max(args.steps, 1), max(args.warmup, 0)
    line = {"metric": metric_name(wl), "value": None, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": None,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "complex128 (f64)", "data": "synthetic",
            "config": workload_config(args, world)}
    t_start = time.perf_counter()
    try:
        kern = CudaKernels(dev)
        ham = Hamiltonian.ising(L, J=1.0, g=0.5) if wl["model"] == "ising" else Hamiltonian.heisenberg(L, 1.0, 1.0, 1.0)
        state = State(tensors=saturated_tensors(L, chi))
        state.mps.normalize()                                 # (the constructor keeps the norm of user tensors, as the reference does)
        noise = NoiseModel(noise_processes(wl))
        observables = [Observable(Z(), s) for s in range(L)] + ([Observable(ZZ(), [s, s + 1]) for s in range(L - 1)] if wl.get("zz") else [])
        n_steps = W + K + 1                                   # time index 0: initial record, 1: initialize + first sample, then W + K steady steps
        sp = AnalogSimParams(observables=observables, elapsed_time=0.1 * n_steps, dt=0.1, num_traj=B * world, max_bond_dim=chi,
                             svd_threshold=wl["threshold"], krylov_tol=1e-4, order=2, sample_timesteps=True, random_seed=42)
        nl = max(1, min(args.lanes, B))
        sim = Simulator(show_progress=False, device=dev, kernels=kern, lanes=nl, on_trajectory_error="flag")
        sim.record_step_events = True
        run = sim.open(state, ham, sp, noise)                 # set-up: H2D of the state, the MPO and the tables' allocation
        assert run.n_times == n_steps + 1, (run.n_times, n_steps)
        run.advance(2 + W)                                    # indices 0, 1 and the W warm-up steps
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        kern.launch_count(reset=True)
        engs = [lane.eng for lane in run.lane_runs]
        stats0 = {k: sum(e.stats[k] for e in engs) for k in engs[0].stats}
        io0 = run.io_bytes()
        instrumented = list(range(max(0, K - max(args.roofline_steps, 0)), K))
        fam_ms: dict = {}
        fam_n: dict = {}
        gemm_flops = 0
        detail: dict = {}
        sweep_events: list = []
        orig_tdvp = {}

        def hook_tdvp(eng):                                   # events around each TDVP sweep of an instrumented step ("TDVP sweep ms")
            orig = eng.tdvp

            def timed_tdvp(st, _o=orig):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                _o(st)
                b.record()
                sweep_events.append((a, b))
            orig_tdvp[id(eng)] = orig
            eng.tdvp = timed_tdvp

        with ClockSampler(local) as clocks:
            ref = torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(dev)
            ref.record()
            t0 = time.perf_counter()
            for k in range(K):
                instr = k in instrumented
                if instr:
                    # ONE lane runs alone under per-launch CUDA events (a timed launch must have the GPU to itself); the events are
                    # switched off when it is through and the other lanes do their part of the step concurrently, as in every other step
                    kern.start_timing()
                    hook_tdvp(engs[0])
                    closed: dict = {}

                    def close_timing():
                        closed["timing"] = kern.stop_timing()
                        engs[0].tdvp = orig_tdvp.pop(id(engs[0]))
                    run.advance(1, serial="first", between=close_timing)
                    timing = closed["timing"]
                    gemm_flops += timing.pop("gemm_flops")
                    for key, v in timing.pop("detail", {}).items():
                        cur = detail.setdefault(key, [0, 0.0])
                        cur[0] += v["launches"]
                        cur[1] += v["ms"]
                    for fam, v in timing.items():
                        fam_ms[fam] = fam_ms.get(fam, 0.0) + v["ms"]
                        fam_n[fam] = fam_n.get(fam, 0) + v["launches"]
                else:
                    run.advance(1)
            torch.cuda.synchronize(dev)
            t_wall = time.perf_counter() - t0
            if world > 1:
                dist.barrier()
        launches = kern.launch_count()
        # device time of the K steps: per step, from the earliest start to the latest end over the lanes' events
        first_timed = 2 + W
        spans: dict = {}
        for lane in run.lane_runs:
            for j, e0, e1 in lane.step_events:
                if j >= first_timed:
                    lo, hi = ref.elapsed_time(e0), ref.elapsed_time(e1)
                    a, b = spans.get(j, (lo, hi))
                    spans[j] = (min(a, lo), max(b, hi))
        step_ms = [spans[j][1] - spans[j][0] for j in sorted(spans)]
        assert len(step_ms) == K, (len(step_ms), K)
        lane0_ms = {j - first_timed: e0.elapsed_time(e1) for j, e0, e1 in run.lane_runs[0].step_events if j >= first_timed}      # the instrumented lane's own steps
        ms_dev = float(sum(step_ms))
        t = torch.tensor([ms_dev, t_wall], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_dev, t_wall = float(t[0].item()), float(t[1].item())
        value = B * world * K / (ms_dev / 1000.0)
        line.update(value=value, ms_per_step=ms_dev / K, gpu_launches=launches, clocks=clocks.summary())
        io1 = run.io_bytes()
        line["e2e"] = {"value": B * world * K / t_wall, "unit": UNIT, "h2d_bytes_per_step": int((io1[0] - io0[0]) / K), "d2h_bytes_per_step": int((io1[1] - io0[1]) / K),
                       "seconds": t_wall, "api": "Simulator.open(State, Hamiltonian, AnalogSimParams, NoiseModel) + AnalogRun.advance(1) per step: the step's "
                       "uniforms go pinned host -> device, its result + diagnostics column device -> pinned host; same K steps as `value` by the wall clock",
                       "setup_h2d_bytes": run.h2d_setup_bytes}
        line["step_ms"] = {"min": min(step_ms), "max": max(step_ms), "instrumented_steps": instrumented}

        # ---- counters, whole-step flop model ----------------------------------------------------------------------------
        stats = {k: sum(e.stats[k] for e in engs) - stats0.get(k, 0) for k in engs[0].stats}
        bond_mean = float(torch.cat([e.psi.bond[:, 1:L].to(torch.float64).reshape(-1) for e in engs]).mean().item())
        prof = torch.stack([e.psi.bond.to(torch.float64).mean(dim=0) for e in engs]).mean(dim=0).cpu().tolist()
        mpo_b = [1.0] + [float(w.shape[3]) for w in ham.mpo.tensors]
        try:
            per = float(len(engs) * K)                        # counters are per lane launch; every trajectory of the lane sees them
            n2, n1 = stats["matvecs_2site"] / per, (stats["matvecs"] - stats["matvecs_2site"]) / per
            fm = flop_model_step(prof, mpo_b, n2, n1, stats["jumps"] / float(B * K))
            ws = {"flops_per_trajectory_step_counters": fm["total"], "breakdown": {k: v for k, v in fm.items() if k != "total"},
                  "note": "SURVEY.md 8(d) model on the measured mean bond profile, matvec and jump counts of the timed steps (matvec counts are per "
                          "lane launch: a trajectory whose Lanczos has converged is masked but still counted, so this is an upper bound)"}
            ws["tflops_per_gpu_counters"] = fm["total"] * value / world / 1e12
            ws["frac_of_fp64_peak_counters"] = ws["tflops_per_gpu_counters"] / FP64_PEAK_TFLOPS
            line["whole_step"] = ws
        except Exception as exc:  # noqa: BLE001 -- reporting only
            line["whole_step"] = {"error": repr(exc)}
        line["counters"] = {**{k: int(v) for k, v in stats.items()}, "mean_bond": bond_mean}

        # ---- roofline of the dominant kernel -----------------------------------------------------------------------------
        try:
            if wl["bound"] == "tensor" and instrumented and fam_ms.get("gemm", 0.0) > 0:
                gemm_ms, gemm_n = fam_ms["gemm"], fam_n["gemm"]
                instr_ms = sum(lane0_ms[k] for k in instrumented)      # the instrumented lane's own step (it runs alone)
                achieved = gemm_flops / (gemm_ms / 1000.0) / 1e12
                traffic = None
                try:
                    with open(os.path.join(ROOT, "profiles", "gemm_traffic.json")) as fh:
                        traffic = json.load(fh).get("dram_bytes_per_trajectory_launch") * engs[0].B
                except Exception:
                    pass
                line["roofline"] = {
                    "kernel": "k_gemm (batched complex128 DMMA contraction: project_site / environment updates / merges / Gram products)",
                    "bound": "tensor", "achieved": achieved, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": achieved / FP64_PEAK_TFLOPS, "traffic": traffic,
                    "peak_source": "measured FP64 DMMA issue peak on this pool (profiles/r01_fp64_peak.json); MEASURED_PEAKS.json carries only HBM and bf16",
                    "launches": gemm_n, "avg_launch_ms": gemm_ms / max(gemm_n, 1), "algorithmic_flops_per_launch": gemm_flops / max(gemm_n, 1),
                    "share_of_step": gemm_ms / instr_ms,
                    "timing": f"CUDA events around every launch of ONE lane ({engs[0].B} trajectories) inside timed steps {instrumented}: that lane runs "
                              f"alone so that a launch has the GPU to itself, the other lanes do their part of the step concurrently afterwards; "
                              f"flops counted on the device by every CTA (8 per complex multiply-add of its tile)",
                    "instrumented_lane_batch": engs[0].B,
                    "other_kernels_ms_per_step": {k: v / len(instrumented) for k, v in fam_ms.items() if k != "gemm"},
                    "gemm_ms_per_step": gemm_ms / len(instrumented)}
            else:                                             # HBM-bound regime: the whole step against the copy bandwidth
                peak, src = 6468.3, "fallback = the figure MEASURED_PEAKS.json held in round 1"
                try:
                    with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
                        peak, src = float(json.load(fh)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
                except Exception:
                    pass
                by = hbm_bytes_step(prof, mpo_b)
                achieved = by * B * K / (ms_dev / 1000.0) / 1e9
                line["roofline"] = {"kernel": "whole step (state + environment streaming: 15 S + 4 E bytes per trajectory-step, SURVEY.md 8d)", "bound": "hbm",
                                    "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None, "peak_source": src,
                                    "algorithmic_bytes_per_trajectory_step": by,
                                    "note": "at chi <= 16 the step is launch/latency-bound, not bandwidth-bound: the fraction is reported as measured"}
            if detail:                                        # where the instrumented step goes: (family, size key, phase), largest first
                rows = sorted(detail.items(), key=lambda kv: -kv[1][1])
                line["kernel_breakdown"] = [{"family": k[0], "key": k[1], "phase": k[2], "launches": v[0], "ms": round(v[1], 2)} for k, v in rows[:40]]
            if sweep_events:
                sw = [a.elapsed_time(b) for a, b in sweep_events]
                line["tdvp_sweep_ms"] = {"per_batch": float(np.mean(sw)), "batch": engs[0].B, "per_trajectory": float(np.mean(sw)) / engs[0].B, "sweeps_timed": len(sw)}
        except Exception as exc:  # noqa: BLE001
            line["roofline"] = {"error": repr(exc)}

        # ---- the one collective + the checks the timed region owes: status words clean, observables finite -------------------
        res = run.result()
        zmean = np.array([res.expectation_values[s][first_timed + K - 1] for s in range(L)])
        line["consistency"] = {"status_nonzero": int(np.count_nonzero(res.status)), "status_values": sorted({int(v) for v in res.status}),
                               "observables_finite": bool(all(np.isfinite(ev).all() for ev in res.expectation_values)),
                               "max_abs_z": float(np.max(np.abs(zmean))), "z_mean_first_sites": [float(x) for x in zmean[:4]],
                               "trajectories": int(len(res.status)), "columns": int(len(res.times))}
        del run, engs
        torch.cuda.empty_cache()

        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            try:
                cores = args.cpu_cores or os.cpu_count() or 1
                a, b = CPU_SAMPLE_LENGTHS
                lengths = [L] if L <= b else [a, b]
                walls = {n: cpu_pool_wall(wl, n, cores) for n in lengths}
                v, t_full = cpu_extrapolate(wl, walls, cores)
                line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": cpu_sample_text(wl, cores),
                                        "pool_wall_s": walls, "full_chain_half_step_s": t_full}
            except Exception as exc:  # noqa: BLE001
                line["cpu_baseline"] = {"error": repr(exc)}
    except Exception as exc:  # noqa: BLE001 -- never lose the line
        import traceback
        line["error"] = repr(exc)
        line["traceback"] = traceback.format_exc()[-1500:]
    line["bench_wall_s"] = time.perf_counter() - t_start
    if rank == 0:
        emit(line)
    if world > 1:
        try:
            dist.destroy_process_group()
        except Exception:
            pass
    return 0 if "error" not in line else 1


_REAL_STDOUT = None


def _claim_stdout() -> None:
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on the first collective),
    so file descriptor 1 is pointed at stderr for the whole run and the line goes to the saved descriptor."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    args = parse_args()
    _claim_stdout()
    digital = args.wl.get("kind") == "digital"
    if args.impl == "reference":
        return run_reference_digital(args) if digital else run_reference_arm(args)
    return run_b200_digital(args) if digital else run_b200_arm(args)


if __name__ == "__main__":
    sys.exit(main())
