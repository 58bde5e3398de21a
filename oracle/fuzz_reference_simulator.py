"""Seeded random runs through the UNMODIFIED reference's ``Simulator.run`` and through ``yaqs_b200.Simulator.run`` on the numpy kernel
twin (build container only): the public call a user makes, end to end -- validation, static-disorder sampling of the noise model from
the seed, observable sorting and un-sorting, trajectory streams, Result aggregation.  Every field of ``Result`` must agree:
expectation values and per-trajectory rows <= 1e-10, identical times and bond diagnostics, the same sampled noise strengths.
TEST INFRASTRUCTURE; draws the oracle's IllPosedWatch would flag are reported, not compared.

    python -m oracle.fuzz_reference_simulator --cases 100 --seed 0
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

ONE = ["lowering", "raising", "pauli_x", "pauli_y", "pauli_z"]
TWO = ["crosstalk_xx", "crosstalk_zy", "crosstalk_yz", "lowering_two"]
OBS = ["x", "y", "z", "zz", "xx", "h", "p0"]


def draw(rng) -> dict:
    L = int(rng.integers(2, 7))
    procs = []
    for _ in range(int(rng.integers(0, 4))):
        name = str(rng.choice(ONE))
        strength = float(np.round(rng.uniform(0.05, 0.6), 3)) if rng.random() < 0.7 else \
            {"distribution": str(rng.choice(["normal", "lognormal", "truncated_normal"])), "mean": float(np.round(rng.uniform(0.1, 0.4), 3)), "std": 0.05}
        sites = list(range(L)) if rng.random() < 0.6 else [int(rng.integers(0, L))]
        procs += [{"name": name, "sites": [s], "strength": strength} for s in sites]
    if rng.random() < 0.3:
        i = int(rng.integers(0, L - 1))
        procs.append({"name": str(rng.choice(TWO)), "sites": [i, i + 1] if rng.random() < 0.7 else [i + 1, i], "strength": float(np.round(rng.uniform(0.05, 0.5), 3))})
    obs = []
    for _ in range(int(rng.integers(1, 6))):
        name = str(rng.choice(OBS))
        if name in ("zz", "xx"):
            i = int(rng.integers(0, L - 1))
            obs.append((name, [i, i + 1]))
        else:
            obs.append((name, int(rng.integers(0, L))))
    pieces = None
    if rng.random() < 0.25:                               # Hamiltonian.piecewise: static pieces switched on the dt grid (hamiltonian.py:176-263)
        pieces = [(float(np.round(rng.uniform(0.3, 1.4), 3)), float(np.round(rng.uniform(0.1, 1.4), 3)), int(rng.integers(1, 3))) for _ in range(int(rng.integers(2, 4)))]
    return dict(pad=(int(rng.choice([2, 4])) if rng.random() < 0.2 else None), get_state=bool(not procs and rng.random() < 0.6), pieces=pieces, L=L, model=str(rng.choice(["ising", "heisenberg"])), J=float(np.round(rng.uniform(0.3, 1.4), 3)), g=float(np.round(rng.uniform(0.1, 1.4), 3)),
                init=str(rng.choice(["zeros", "ones", "x+", "x-", "y+", "Neel", "wall"])), procs=procs, obs=obs,
                params=dict(elapsed_time=0.1 * int(rng.integers(1, 5)), dt=0.1, num_traj=int(rng.integers(1, 6)), max_bond_dim=int(rng.choice([2, 4, 8, 16])),
                            svd_threshold=float(10.0 ** rng.uniform(-11, -5)), order=int(rng.choice([1, 2])), sample_timesteps=bool(rng.random() < 0.7),
                            random_seed=int(rng.integers(0, 10000)), tdvp_sweeps=int(rng.choice([1, 1, 2, 3])),
                            trunc_mode=str(rng.choice(["discarded_weight", "discarded_weight", "relative", "hard_cutoff", "relative_discarded_weight"])),
                            krylov_tol=float(rng.choice([1e-4, 1e-6, 1e-10])), preset=str(rng.choice(["fast", "balanced", "accurate", "exact"]))))


def build(ns, cfg):
    def static(j, g):
        return ns["Hamiltonian"].ising(cfg["L"], j, g) if cfg["model"] == "ising" else ns["Hamiltonian"].heisenberg(cfg["L"], j, j, g)
    params_kw = dict(cfg["params"])
    if cfg.get("pieces"):
        ham = ns["Hamiltonian"].piecewise([(static(j, g), 0.1 * n) for j, g, n in cfg["pieces"]])
        params_kw["elapsed_time"] = 0.1 * sum(n for _, _, n in cfg["pieces"])
    else:
        ham = static(cfg["J"], cfg["g"])
    obs = [ns["Observable"](name, sites) for name, sites in cfg["obs"]]
    noise = ns["NoiseModel"]([dict(p) for p in cfg["procs"]]) if cfg["procs"] else None
    params = ns["AnalogSimParams"](observables=obs, get_state=bool(cfg.get("get_state")), **params_kw)
    return ns["State"](cfg["L"], initial=cfg["init"], pad=cfg.get("pad")), ham, params, noise


def compare(cfg, ref_ns, our_ns, ref_sim, our_sim):
    from fuzz_twin_vs_oracle import IllPosedWatch
    import mqt.yaqs.core.linalg as ref_linalg
    import oracle.tjm_oracle as O
    orig = ref_linalg.truncate

    def watched(s, *, mode, threshold, max_bond_dim=None, min_keep=1):
        O.truncate(np.asarray(s), mode, threshold, max_bond_dim, min_keep)        # the watch sits on the oracle's truncate
        return orig(s, mode=mode, threshold=threshold, max_bond_dim=max_bond_dim, min_keep=min_keep)

    with IllPosedWatch() as watch:                        # every truncation decision of the reference run goes past the watch
        ref_linalg.truncate = watched
        try:
            ref = ref_sim.run(*build(ref_ns, cfg))
        finally:
            ref_linalg.truncate = orig
    ours = our_sim.run(*build(our_ns, cfg))
    why, worst = None, 0.0
    if not np.array_equal(np.asarray(ref.times), np.asarray(ours.times)):
        why = "times differ"
    for k in range(len(cfg["obs"])):
        if why:
            break
        a, b = np.asarray(ref.expectation_values[k]), np.asarray(ours.expectation_values[k])
        ta, tb = np.asarray(ref.trajectories[k]), np.asarray(ours.trajectories[k])
        if a.shape != b.shape or ta.shape != tb.shape:
            why = f"observable {k}: shapes {a.shape} / {ta.shape} against {b.shape} / {tb.shape}"
            break
        worst = max(worst, float(np.max(np.abs(ta - tb))), float(np.max(np.abs(a - b))))
        if not worst < 1e-10:
            why = f"observable {k} differs by {worst:.3e}"
    if why is None:
        for name in ("runtime_cost", "max_bond", "total_bond"):
            if not np.array_equal(np.asarray(getattr(ref, name)), np.asarray(getattr(ours, name))):
                why = f"{name} differs: {getattr(ref, name)} against {getattr(ours, name)}"
                break
    if why is None and cfg.get("get_state"):
        va, vb = ref.output_state.mps.to_vec(), ours.output_state.mps.to_vec()
        if abs(abs(np.vdot(va, vb)) - np.linalg.norm(va) * np.linalg.norm(vb)) > 1e-10 or abs(np.linalg.norm(va) - np.linalg.norm(vb)) > 1e-10:
            why = f"output_state differs: overlap {abs(np.vdot(va, vb))}, norms {np.linalg.norm(va)} / {np.linalg.norm(vb)}"
    if why is None and cfg["procs"]:
        sa = [p["strength"] for p in ref.noise_model.processes]
        sb = [p["strength"] for p in ours.noise_model.processes]
        if sa != sb:
            why = "sampled noise strengths differ"
    if why is not None and watch.notes:
        return "ill-posed: " + watch.notes[0] + " [" + why + "]", worst
    return why, worst


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
    from oracle.np_kernels import NumpyKernels
    ref_ns = dict(Hamiltonian=RH, NoiseModel=RN, AnalogSimParams=RA, Observable=RO, State=RS)
    our_ns = dict(Hamiltonian=Y.Hamiltonian, NoiseModel=Y.NoiseModel, AnalogSimParams=Y.AnalogSimParams, Observable=Y.Observable, State=Y.State)
    ref_sim = ref_simulator.Simulator(parallel=False, show_progress=False)
    our_sim = Y.Simulator(show_progress=False, kernels=NumpyKernels())
    rng = np.random.default_rng(args.seed)
    bad = 0
    for n in range(args.cases):
        cfg = draw(rng)
        t0 = time.time()
        try:
            why, worst = compare(cfg, ref_ns, our_ns, ref_sim, our_sim)
        except Exception as exc:  # noqa: BLE001
            why, worst = f"{type(exc).__name__}: {exc}", float("nan")
        mark = "ok " if why is None else ("-- " if why.startswith("ill-posed") else "BAD")
        print(f"[{n:3d}] {mark} worst {worst:.1e} {time.time() - t0:5.1f}s  {cfg if why else ''}{' -> ' + why if why else ''}", flush=True)
        if why is not None and not why.startswith("ill-posed"):
            bad += 1
            if not args.keep_going:
                return 1
    print(f"{args.cases} cases, {bad} mismatches")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
