"""Seeded random circuits through the UNMODIFIED reference's digital backend (``digital_tjm`` per trajectory, hand-compiled layers because
qiskit is absent; trajectory generators from the seed as the reference's worker makes them) and through ``yaqs_b200.Simulator.run(State,
Circuit, DigitalSimParams, NoiseModel)`` on the numpy kernel twin: the public digital call end to end -- circuit compilation, gate
routes, per-gate local noise, sampling barriers, Result aggregation.  Per-trajectory rows and means <= 1e-10.  TEST INFRASTRUCTURE, build
container only.

    python -m oracle.fuzz_reference_digital_simulator --cases 100 --seed 0
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


def compare(cfg: dict):
    from fuzz_twin_vs_oracle import IllPosedWatch
    import mqt.yaqs.core.linalg as ref_linalg
    import mqt.yaqs.digital.digital_tjm as ref_dig
    import oracle.tjm_oracle as O
    import yaqs_b200 as Y
    from mqt.yaqs.core.data_structures.mps import MPS as RMPS
    from mqt.yaqs.core.data_structures.noise_model import NoiseModel as RN
    from mqt.yaqs.core.data_structures.simulation_parameters import DigitalSimParams as RD
    from mqt.yaqs.core.data_structures.simulation_parameters import Observable as RO
    from mqt.yaqs.core.libraries.gate_library import GateLibrary as RG
    from oracle.np_kernels import NumpyKernels
    from yaqs_b200.circuits import SAMPLE_LABEL, Circuit
    L = cfg["L"]
    layers, circ = [], Circuit(L)
    for name, params, qubits in cfg["gates"]:
        if name == "sample":
            layers.append(ref_dig._CompiledCircuitLayer((), (), (), 1))
            circ.barrier(label=SAMPLE_LABEL)
            continue
        gate = getattr(RG, name)(list(params)) if params else getattr(RG, name)()
        gate.set_sites(*qubits)
        if len(qubits) == 1:
            layers.append(ref_dig._CompiledCircuitLayer((gate,), (), (), 0))
        elif min(qubits) % 2 == 0:
            layers.append(ref_dig._CompiledCircuitLayer((), (gate,), (), 0))
        else:
            layers.append(ref_dig._CompiledCircuitLayer((), (), (gate,), 0))
        getattr(circ, name)(*params, *qubits)
        circ.barrier()
    compiled = ref_dig._CompiledCircuit(tuple(layers), sum(l.sample_points for l in layers))
    obs_spec = [("z", s) for s in range(L)] + ([("x", s) for s in range(L)] if "x" in cfg["obs"] else [])
    procs = [{"name": n, "sites": [i], "strength": g} for i in range(L) for (n, g) in cfg["noise"]]
    common = dict(num_traj=cfg["ntraj"], max_bond_dim=cfg["chi"], svd_threshold=cfg["thr"], sample_layers=cfg["sample"], random_seed=cfg["seed"],
                  gate_mode=cfg["gate_mode"])
    sim_r = RD(observables=[RO(n, s) for n, s in obs_spec], **common)
    sim_r.num_mid_measurements = compiled.num_mid_measurements
    noise_r = RN([dict(p) for p in procs]) if procs else None
    ntraj = cfg["ntraj"] if procs else 1
    orig = ref_linalg.truncate

    def watched(s, *, mode, threshold, max_bond_dim=None, min_keep=1):
        O.truncate(np.asarray(s), mode, threshold, max_bond_dim, min_keep)
        return orig(s, mode=mode, threshold=threshold, max_bond_dim=max_bond_dim, min_keep=min_keep)
    rows = []
    with IllPosedWatch() as watch:
        ref_linalg.truncate = watched
        try:
            for t in range(ntraj):
                res, _diag, _counts, _final = ref_dig.digital_tjm((t, RMPS(L, state=cfg["init"]), noise_r, sim_r, None), compiled_circuit=compiled)
                rows.append(np.asarray(res, dtype=np.float64))
        finally:
            ref_linalg.truncate = orig
    rows = np.array(rows)                                                       # [traj, sorted obs, cols]
    sp = Y.DigitalSimParams(observables=[Y.Observable(n, s) for n, s in obs_spec], **common)
    ours = Y.Simulator(show_progress=False, kernels=NumpyKernels()).run(Y.State(L, initial=cfg["init"]), circ, sp, Y.NoiseModel(procs) if procs else None)
    order = np.argsort([s for _, s in obs_spec], kind="stable")                # the reference's workers see the observables sorted by site
    why, worst = None, 0.0
    for user_i, sorted_i in enumerate(np.argsort(order, kind="stable")):
        a = rows[:, sorted_i, :]
        b = np.asarray(ours.trajectories[user_i])
        if a.shape != b.shape:
            why = f"observable {user_i}: shapes {a.shape} against {b.shape}"
            break
        worst = max(worst, float(np.max(np.abs(a - b))), float(np.max(np.abs(a.mean(axis=0) - np.asarray(ours.expectation_values[user_i])))))
        if not worst < 1e-10:
            why = f"observable {user_i} differs by {worst:.3e}"
            break
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
    from fuzz_digital_twin_vs_oracle import draw_gate_case
    rng = np.random.default_rng(args.seed)
    bad = 0
    for n in range(args.cases):
        cfg = draw_gate_case(rng)
        cfg["init"] = str(rng.choice(["zeros", "x+", "Neel"]))                  # the public call takes preset states
        if cfg["gate_mode"] in ("tdvp", "full-tdvp"):
            cfg["gate_mode"] = str(rng.choice(["mpo", "swaps"]))                # (TDVP windows on product states are ill-posed: min_keep = 2)
        t0 = time.time()
        try:
            why, worst = compare(cfg)
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
