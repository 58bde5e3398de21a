"""Seeded random analog SimulationPrograms through the UNMODIFIED reference's ``Simulator.run(state, segments, None, noise, observables=...,
num_traj=..., random_seed=...)`` and through ``yaqs_b200.Simulator.run`` on the numpy kernel twin (build container only): the stitched
outer Result (expectation values, times) and every segment Result (trajectory rows, bond diagnostics, time offsets).  Digital segments
need qiskit in the reference's public call and are covered at the executor level (oracle/fuzz_reference_vs_oracle.py --program).
TEST INFRASTRUCTURE.

    python -m oracle.fuzz_reference_program_simulator --cases 100 --seed 0
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np
import torch  # noqa: F401 -- before the shim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "walks"))
for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
    os.environ.setdefault(_v, "1")


def build(ns, cfg):
    L = cfg["L"]
    obs = []
    for tok in cfg["obs"].split("+"):
        obs += [ns["Observable"](tok, [s, s + 1]) for s in range(L - 1)] if tok == "zz" else [ns["Observable"](tok, s) for s in range(L)]
    procs = [{"name": n, "sites": [i], "strength": g} for i in range(L) for (n, g) in cfg["noise"]]
    noise = ns["NoiseModel"](procs) if procs else None
    hams, segs = {}, []
    for seg in cfg["segments"]:
        if seg["ham"] not in hams:
            hams[seg["ham"]] = ns["Hamiltonian"].ising(L, seg["J"], seg["g"])
        segs.append((hams[seg["ham"]], ns["AnalogSimParams"](elapsed_time=seg["T"], dt=seg["dt"], max_bond_dim=seg["chi"], svd_threshold=seg["thr"],
                                                            order=seg["order"], sample_timesteps=seg["sample"])))
    return ns["State"](L, initial=cfg["init"]), segs, noise, obs


def compare(cfg, ref_ns, our_ns, ref_sim, our_sim):
    state, segs, noise, obs = build(ref_ns, cfg)
    ref = ref_sim.run(state, segs, None, noise, observables=obs, num_traj=cfg["ntraj"], random_seed=cfg["seed"])
    state, segs, noise, obs = build(our_ns, cfg)
    ours = our_sim.run(state, segs, None, noise, observables=obs, num_traj=cfg["ntraj"], random_seed=cfg["seed"])
    worst = 0.0
    if not np.allclose(np.asarray(ref.times), np.asarray(ours.times), rtol=0, atol=1e-12):
        return f"times differ: {ref.times} against {ours.times}", worst
    if len(ref.segment_results) != len(ours.segment_results):
        return f"{len(ours.segment_results)} segment results, the reference has {len(ref.segment_results)}", worst
    for k in range(len(obs)):
        a, b = np.asarray(ref.expectation_values[k]), np.asarray(ours.expectation_values[k])
        if a.shape != b.shape:
            return f"observable {k}: stitched shapes {a.shape} against {b.shape}", worst
        worst = max(worst, float(np.max(np.abs(a - b))))
    for s, (ra, rb) in enumerate(zip(ref.segment_results, ours.segment_results)):
        for k in range(len(obs)):
            ta, tb = np.asarray(ra.trajectories[k]), np.asarray(rb.trajectories[k])
            if ta.shape != tb.shape:
                return f"segment {s}, observable {k}: shapes {ta.shape} against {tb.shape}", worst
            worst = max(worst, float(np.max(np.abs(ta - tb))))
        for name in ("runtime_cost", "max_bond", "total_bond"):
            if not np.array_equal(np.asarray(getattr(ra, name)), np.asarray(getattr(rb, name))):
                return f"segment {s}: {name} differs", worst
        if ra.segment_index != rb.segment_index or abs(float(ra.time_offset) - float(rb.time_offset)) > 1e-12 or ra.segment_type != rb.segment_type:
            return f"segment {s}: bookkeeping differs", worst
    return (None if worst < 1e-10 else f"observables differ by {worst:.3e}"), worst


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=60)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--keep-going", action="store_true")
    args = ap.parse_args()
    from oracle.ref_shim import load_reference
    load_reference()
    import mqt.yaqs.simulator as ref_simulator
    from mqt.yaqs.core.data_structures.hamiltonian import Hamiltonian as RH
    from mqt.yaqs.core.data_structures.noise_model import NoiseModel as RN
    from mqt.yaqs.core.data_structures.simulation_parameters import AnalogSimParams as RA
    from mqt.yaqs.core.data_structures.simulation_parameters import Observable as RO
    from mqt.yaqs.core.data_structures.state import State as RS
    import yaqs_b200 as Y
    from fuzz_program_twin_vs_oracle import draw_case
    from oracle.np_kernels import NumpyKernels
    ref_ns = dict(Hamiltonian=RH, NoiseModel=RN, AnalogSimParams=RA, Observable=RO, State=RS)
    our_ns = dict(Hamiltonian=Y.Hamiltonian, NoiseModel=Y.NoiseModel, AnalogSimParams=Y.AnalogSimParams, Observable=Y.Observable, State=Y.State)
    ref_sim = ref_simulator.Simulator(parallel=False, show_progress=False)
    our_sim = Y.Simulator(show_progress=False, kernels=NumpyKernels())
    rng = np.random.default_rng(args.seed)
    bad = 0
    for n in range(args.cases):
        cfg = draw_case(rng)
        cfg["segments"] = [s for s in cfg["segments"] if s["kind"] == "A"] or [dict(kind="A", ham="h0", J=1.0, g=0.5, T=0.2, dt=0.1, order=1, chi=8, thr=1e-9, sample=True)]
        t0 = time.time()
        try:
            why, worst = compare(cfg, ref_ns, our_ns, ref_sim, our_sim)
        except Exception as exc:  # noqa: BLE001
            why, worst = f"{type(exc).__name__}: {exc}", float("nan")
        print(f"[{n:3d}] {'ok ' if why is None else 'BAD'} worst {worst:.1e} {time.time() - t0:5.1f}s  {cfg if why else ''}{' -> ' + why if why else ''}", flush=True)
        if why is not None:
            bad += 1
            if not args.keep_going:
                return 1
    print(f"{args.cases} cases, {bad} mismatches")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
