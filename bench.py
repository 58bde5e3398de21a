#!/usr/bin/env python
"""bench.py -- noisy-MPS TJM throughput on B200 (BASELINE.json metric) with roofline and CPU baseline.

    python bench.py --gpus N --steps K --warmup W              # our arm (torchrun launches one rank per GPU for N > 1)
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU arm: the oracle port on the host cores

Workload (config.workload = "C3"): BASELINE.json configs[2] -- transverse-field Ising chain L=100, chi_max=128,
relaxation+dephasing (lowering + pauli_z, gamma=0.1 on every site), second-order TJM with observables (Z on every site)
sampled every step, started from a seeded random MPS whose bonds are saturated at chi (otherwise the first
~50 steps run at tiny chi and measure nothing).  A "step" is one steady-state sampled TJM time step of one batch of
`--batch` trajectories per GPU: step_through (TDVP sweep, dissipation, jump) + sample (copy, TDVP sweep, dissipation,
jump, diagnostics, observables); the K timed steps end with the one all-reduce of the per-time-step observable sums.  Weak scaling: the batch per GPU is
fixed, value = (batch * N * K) / max-over-ranks device time.
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

METRIC = "noisy-MPS trajectory-steps/s (Ising L=100, chi=128, order-2 TJM, sampled every step)"
UNIT = "trajectory-steps/s"
FP64_PEAK_TFLOPS = 37.1      # measured on this pool's B200: profiles/r01_fp64_peak.json (DMMA == DFMA issue peak)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=int(os.environ.get("YB_BENCH_BATCH", "256")), help="trajectories per GPU per step")
    ap.add_argument("--lanes", type=int, default=int(os.environ.get("YB_BENCH_LANES", "2")),
                    help="concurrent lock-step sub-batches per GPU (host thread + stream each)")
    ap.add_argument("--length", type=int, default=100)
    ap.add_argument("--chi", type=int, default=128)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-roofline", action="store_true", help="skip the instrumented extra step and the sweep timing (A/B runs)")
    ap.add_argument("--cpu-cores", type=int, default=0, help="cores for the CPU arm (0 = all)")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------------------------------
# synthetic workload (host side)
# ----------------------------------------------------------------------------------------------------------------------
def saturated_tensors(length: int, chi: int, seed: int = 1234):
    """Seeded random MPS with bonds min(chi, 2**k); normalised by MPS(tensors=...) / the oracle's normalize."""
    rng = np.random.default_rng(seed)
    bonds = [1] + [min(chi, 2 ** min(i + 1, length - 1 - i)) for i in range(length - 1)] + [1]
    return [(rng.standard_normal((2, bonds[i], bonds[i + 1])) + 1j * rng.standard_normal((2, bonds[i], bonds[i + 1]))).astype(np.complex128)
            for i in range(length)]


def workload_config(args, nranks):
    return {"workload": "C3", "model": None, "hamiltonian": "ising J=1 g=0.5", "length": args.length, "chi_max": args.chi,
            "noise": "lowering+pauli_z gamma=0.1 per site", "order": 2, "dt": 0.1, "svd_threshold": 1e-6, "krylov_tol": 1e-4,
            "observables": "Z on every site, every step", "initial_state": "seeded random MPS, bonds saturated at chi",
            "batch_per_gpu": args.batch, "global_batch": args.batch * nranks, "sharding": f"trajectories x{nranks}", "lanes_per_gpu": args.lanes,
            "l2_policy": "inputs larger than L2 (per-step working set >> 126 MB)",
            "library_options": {k: os.environ[k] for k in ("YB_GEMM_RAGGED", "YB_WAPPLY_ONEPASS", "YB_HEEV_RING", "YB_HEEV_QL") if k in os.environ}}


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
            self._stop.wait(0.25)

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
def _cpu_half_step(job):
    """One step_through (TDVP sweep + dissipation(dt) + jump test) of ONE trajectory = half a sampled order-2 step."""
    length, chi, traj = job
    from threadpoolctl import threadpool_limits

    from oracle import tjm_oracle as O
    with threadpool_limits(1):
        state = O.OracleMPS(saturated_tensors(length, chi), None)
        O.normalize(state, "B", "QR")
        p = O.OracleParams(dt=0.1, n_steps=1, order=2, max_bond_dim=chi, svd_threshold=1e-6, krylov_tol=1e-4)
        mpo = O.ising_mpo(length, 1.0, 0.5)
        noise = [O.make_process(n, [i], 0.1) for i in range(length) for n in ("lowering", "pauli_z")]
        rng = O.UniformStream(O.trajectory_stream(42, traj))
        t0 = time.perf_counter()
        O.tdvp(state, mpo, p)
        O.apply_dissipation(state, noise, p.dt, p)
        O.stochastic_process(state, noise, p.dt, p, rng)
        return time.perf_counter() - t0


def cpu_sample(length: int, chi: int, cores: int):
    """All `cores` workers run one half-step each, concurrently (the reference's executor is a fork pool with one BLAS
    thread per worker, core/parallel_utils.py:307-358).  Returns (trajectory-steps/s, wall seconds)."""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        per = pool.map(_cpu_half_step, [(length, chi, t) for t in range(cores)])
    wall = time.perf_counter() - t0
    return 0.5 * cores / wall, wall, per


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = args.cpu_cores or os.cpu_count() or 1
    for _ in range(args.warmup):
        cpu_sample(10, 16, min(cores, 2))                       # warm the pool / BLAS on a tiny chain (untimed)
    t0 = time.perf_counter()
    vals = []
    for _ in range(args.steps):
        v, wall, _ = cpu_sample(args.length, args.chi, cores)
        vals.append(v)
    total = time.perf_counter() - t0
    value = float(np.mean(vals))
    sample = (f"{cores} trajectories in parallel (1 BLAS thread each), each one step_through = TDVP sweep + dissipation(dt) + jump test "
              f"at L={args.length}, chi={args.chi} from the saturated state = 0.5 sampled order-2 step; linear in steps")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 * total / max(args.steps, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "complex128 (f64)", "data": "synthetic", "config": workload_config(args, 1),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------------------
def run_b200_arm(args):
    import torch
    import torch.distributed as dist

    from yaqs_b200 import AnalogSimParams, Hamiltonian, NoiseModel, Observable, Simulator, State, Z
    from yaqs_b200.engine import BatchedTJM, EngineParams
    from yaqs_b200.kernels import CudaKernels
    from yaqs_b200.rng import numpy_tables

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L, chi, B, K, W = args.length, args.chi, args.batch, args.steps, max(args.warmup, 0)
    kern = CudaKernels(dev)
    ham = Hamiltonian.ising(L, J=1.0, g=0.5)
    state = State(tensors=saturated_tensors(L, chi))
    noise = NoiseModel([{"name": n, "sites": [i], "strength": 0.1} for i in range(L) for n in ("lowering", "pauli_z")])
    obs = [(np.array([[1, 0], [0, -1]], dtype=np.complex128), [s]) for s in range(L)]
    ep = EngineParams(dt=0.1, n_steps=K + W + 2, order=2, max_bond_dim=chi, svd_threshold=1e-6, krylov_tol=1e-4)
    from yaqs_b200.lanes import Lanes
    nl = max(1, min(args.lanes, B))
    lanes = Lanes(dev, nl)
    bounds = [(k * B) // nl for k in range(nl + 1)]
    engs = [BatchedTJM(kern, L, bounds[k + 1] - bounds[k], ham.mpo.tensors, ep, noise.processes, obs, state.mps.tensors) for k in range(nl)]
    traj_ids = [rank * B + b for b in range(B)]
    tu, su = numpy_tables(42, traj_ids, ep.n_times, 2)
    tu_d, su_d = torch.from_numpy(tu).to(dev), torch.from_numpy(su).to(dev)

    def lane_steps(job):
        """Lane k advances `count` steady-state steps back to back (no cross-lane barrier inside: lanes are independent
        trajectories, exactly as in Simulator.run); returns the per-step sums [count, L + 3]."""
        k, j0, count = job
        sl = slice(bounds[k], bounds[k + 1])
        rows = []
        for j in range(j0, j0 + count):
            o, dg = engs[k].steady_state_step(tu_d[sl, 2 * j: 2 * j + 2], su_d[sl, j])
            rows.append(torch.cat([o.sum(dim=0), dg.sum(dim=0)]))
        return torch.stack(rows)

    def run_steps(j0, count, concurrent=True):
        jobs = [(k, j0, count) for k in range(nl)]
        parts = lanes.map(lane_steps, jobs) if concurrent else [lane_steps(job) for job in jobs]
        sums = torch.zeros((count, L + 4), dtype=torch.float64, device=dev)
        sums[:, :L + 3] = torch.stack(parts).sum(dim=0)
        sums[:, L + 3] = float(B)
        if world > 1:
            dist.all_reduce(sums)                       # the one collective of the path: per-time-step observable sums + counts
        return sums

    # initialize (F0): half-step dissipation + jump test, then warm-up steps
    def lane_init(k):
        sl = slice(bounds[k], bounds[k + 1])
        engs[k].apply_dissipation(engs[k].phi, ep.dt / 2)
        engs[k].stochastic_process(engs[k].phi, ep.dt, tu_d[sl, 0].contiguous(), tu_d[sl, 1].contiguous())

    lanes.map(lane_init, range(nl))
    if W > 0:
        run_steps(1, W)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    kern.launch_count(reset=True)
    stats0 = {k: sum(e.stats[k] for e in engs) for k in engs[0].stats}
    with ClockSampler(local) as clocks:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        e0.record()
        sums = run_steps(W + 1, K)
        e1.record()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        ms_total = e0.elapsed_time(e1)
    launches = kern.launch_count()
    obs_mean = (sums[-1, :L] / sums[-1, L + 3]).cpu().numpy()
    stats = {k: sum(e.stats[k] for e in engs) - stats0.get(k, 0) for k in engs[0].stats}
    bond_mean = float(torch.cat([e.psi.bond[:, 1:L].to(torch.float64).reshape(-1) for e in engs]).mean().item())
    whole_step = None
    try:                                                     # reporting only: never let the model break the bench line
        prof = torch.stack([e.psi.bond.to(torch.float64).mean(dim=0) for e in engs]).mean(dim=0).cpu().tolist()
        mpo_b = [1.0] + [float(w.shape[3]) for w in ham.mpo.tensors]
        per = float(nl * K)                                   # counters are per lane launch; every trajectory of the lane sees them
        n2, n1 = stats["matvecs_2site"] / per, (stats["matvecs"] - stats["matvecs_2site"]) / per
        fm = flop_model_step(prof, mpo_b, n2, n1, stats["jumps"] / float(B * K))
        whole_step = {"flops_per_trajectory_step_counters": fm["total"], "breakdown": {k: v for k, v in fm.items() if k != "total"},
                      "flops_per_trajectory_step_survey": 2.15e12,
                      "note": "SURVEY.md 8(d) model; 'counters' = measured mean bond profile, matvec and jump counts of the timed steps; "
                              "'survey' = the headline figure for uniform chi = 128.  Matvec counts are per lane launch (a trajectory whose Lanczos has "
                              "converged is masked but still counted), so 'counters' is an upper bound on the algorithmic flops"}
    except Exception as exc:  # noqa: BLE001
        whole_step = {"error": repr(exc)}
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = B * world * K / (ms_total / 1000.0)
    if whole_step and "error" not in whole_step:
        for key in ("counters", "survey"):
            tf = whole_step[f"flops_per_trajectory_step_{key}"] * value / world / 1e12
            whole_step[f"tflops_per_gpu_{key}"] = tf
            whole_step[f"frac_of_fp64_peak_{key}"] = tf / FP64_PEAK_TFLOPS

    roofline, sweep_ms, sweep_batch = None, float("nan"), engs[0].B
    if not args.no_roofline:
        # instrumented pass for the roofline: ONE more step of the same kind, lanes run one after the other so that the CUDA
        # events around every launch see that kernel alone on the GPU (under concurrent lanes they would overlap)
        kern.start_timing()
        i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        i0.record()
        run_steps(W + 1 + K, 1, concurrent=False)
        i1.record()
        torch.cuda.synchronize(dev)
        ms_instr = i0.elapsed_time(i1)
        timing = kern.stop_timing()
        eng = engs[0]

        # TDVP sweep alone (BASELINE.json: "TDVP sweep ms")
        torch.cuda.synchronize(dev)
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        eng.psi.copy_from(eng.phi)
        s0.record()
        eng.tdvp(eng.psi)
        s1.record()
        torch.cuda.synchronize(dev)
        sweep_ms = s0.elapsed_time(s1)
        sweep_batch = eng.B

        gemm = timing.get("gemm", {"launches": 0, "ms": 0.0})
        achieved = timing["gemm_flops"] / (gemm["ms"] / 1000.0) / 1e12 if gemm["ms"] > 0 else 0.0
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "gemm_traffic.json")) as fh:
                traffic = json.load(fh).get("dram_bytes_per_trajectory_launch") * (bounds[1] - bounds[0])     # ncu capture scaled to this lane batch
        except Exception:
            pass
        roofline = {"kernel": "k_gemm (batched complex128 DMMA contraction: project_site / environment updates / merges)",
                    "bound": "tensor", "achieved": achieved, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": achieved / FP64_PEAK_TFLOPS,
                    "traffic": traffic, "peak_source": "measured FP64 DMMA issue peak on this pool (profiles/r01_fp64_peak.json); MEASURED_PEAKS.json "
                    "carries only HBM and bf16", "launches": gemm["launches"], "avg_launch_ms": gemm["ms"] / max(gemm["launches"], 1),
                    "algorithmic_flops": timing["gemm_flops"], "share_of_step": gemm["ms"] / ms_instr,
                    "timing": f"instrumented extra step, lanes serialised ({ms_instr:.0f} ms vs {ms_total / K:.0f} ms per step with {nl} concurrent lanes)",
                    "other_kernels_ms": {k: v["ms"] for k, v in timing.items() if isinstance(v, dict) and k != "gemm"}}


    # free the resident engine before the end-to-end pass
    eng = None
    del engs
    torch.cuda.empty_cache()

    e2e = None
    if not args.no_e2e:
        sp = AnalogSimParams(observables=[Observable(Z(), s) for s in range(L)], elapsed_time=0.1 * K, dt=0.1, num_traj=B * world, max_bond_dim=chi,
                             svd_threshold=1e-6, krylov_tol=1e-4, order=2, sample_timesteps=True, random_seed=42)
        sim = Simulator(show_progress=False, device=dev, kernels=kern, lanes=nl)
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        res = sim.run(state, ham, sp, noise)
        torch.cuda.synchronize(dev)
        t_e2e = time.perf_counter() - t0
        te = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        t_e2e = float(te.item())
        h2d = sum(x.nbytes for x in state.mps.tensors) * 1 + B * (2 * (K + 1) + 2 * (K + 1)) * 8 + sum(w.nbytes for w in ham.mpo.tensors) * 3
        d2h = B * (L + 3) * (K + 1) * 8
        e2e = {"value": B * world * K / t_e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d / K), "d2h_bytes_per_step": int(d2h / K),
               "seconds": t_e2e, "api": "Simulator.run(State, Hamiltonian, AnalogSimParams, NoiseModel) with host inputs; includes buffer "
               "allocation, H2D of state/MPO/uniform tables, D2H of per-trajectory results", "z_mean_site0_last": float(res.expectation_values[0][-1]),
               "steps_covered": f"initialisation + steps 1..{K} from the initial state (value: steps {W + 1}..{W + K} of the resident batch)",
               "counters": {k: int(v) for k, v in (res.stats or {}).items()}}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = args.cpu_cores or os.cpu_count() or 1
        v, wall, per = cpu_sample(L, chi, cores)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "wall_s": wall, "single_core_half_step_s": float(np.mean(per)),
               "sample": f"{cores} trajectories in parallel (fork pool, 1 BLAS thread each), each one step_through (TDVP sweep + dissipation(dt) + "
                         f"jump test) at L={L}, chi={chi} from the saturated state = 0.5 sampled order-2 step"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_total / K,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "complex128 (f64)", "data": "synthetic",
                "config": workload_config(args, world), "roofline": roofline, "whole_step": whole_step, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks.summary(),
                "gpu_launches": launches, "tdvp_sweep_ms": {"per_batch": sweep_ms, "batch": sweep_batch, "per_trajectory": sweep_ms / sweep_batch},
                "counters": {**stats, "mean_bond": bond_mean, "z_mean_first_sites": [float(x) for x in obs_mean[:4]]}}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_b200_arm(args)


if __name__ == "__main__":
    sys.exit(main())
