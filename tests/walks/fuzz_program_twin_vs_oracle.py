"""Seeded random SimulationPrograms: ``Simulator.run`` on the numpy kernel twin against the oracle's program executor.
TEST INFRASTRUCTURE (CPU only), companion of tests/walks/fuzz_twin_vs_oracle.py.

Draws segment lists (analog order 1 / order 2 with shared or different Hamiltonians, equal or different settings -- so that merges,
order-2 phi hand-offs, sample-stream offsets and final-only sampling all occur -- and digital Trotter blocks in between) and asks, per
segment and trajectory, for the same observables (<= 1e-10), identical bond diagnostics, and the same number of doubles drawn from
each trajectory's shared stream.

    python tests/walks/fuzz_program_twin_vs_oracle.py --cases 40 --seed 0
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
    os.environ.setdefault(_v, "1")


def draw_case(rng: np.random.Generator) -> dict:
    L = int(rng.integers(3, 7))
    hams = {f"h{k}": (float(np.round(rng.uniform(0.4, 1.4), 3)), float(np.round(rng.uniform(0.2, 1.5), 3))) for k in range(2)}
    settings = [dict(chi=int(rng.choice([4, 8])), thr=float(10.0 ** rng.uniform(-11, -7))) for _ in range(2)]
    nseg = int(rng.integers(1, 5))
    segs = []
    order_bias = int(rng.choice([1, 2]))
    sample = bool(rng.random() < 0.6)
    for _ in range(nseg):
        if rng.random() < 0.2:
            segs.append(dict(kind="D", J=float(np.round(rng.uniform(0.4, 1.4), 3)), g=float(np.round(rng.uniform(0.2, 1.2), 3)), dt=0.1,
                             timesteps=int(rng.integers(1, 3)), chi=8, thr=1e-9, sample=bool(rng.random() < 0.5)))
            continue
        name = str(rng.choice(list(hams)))
        st = settings[int(rng.random() < 0.3)]
        order = order_bias if rng.random() < 0.8 else 3 - order_bias
        segs.append(dict(kind="A", ham=name, J=hams[name][0], g=hams[name][1], T=0.1 * int(rng.integers(1, 4)), dt=0.1, order=order,
                         chi=st["chi"], thr=st["thr"], sample=sample if rng.random() < 0.85 else not sample))
    noise = [(str(n), float(np.round(rng.uniform(0.05, 0.6), 3))) for n in rng.choice(["lowering", "pauli_z", "pauli_x", "raising"], size=int(rng.integers(0, 3)), replace=False)]
    return dict(L=L, init=str(rng.choice(["zeros", "x+", "Neel"])), noise=noise, ntraj=int(rng.integers(1, 4)), seed=int(rng.integers(0, 10000)),
                obs=str(rng.choice(["z", "z+x", "z+zz"])), segments=segs)


def run_one(cfg: dict):
    import yaqs_b200 as Y
    from fuzz_twin_vs_oracle import IllPosedWatch
    from oracle import tjm_oracle as O
    from oracle.golden_cases import program_oracle_inputs
    from oracle.np_kernels import NumpyKernels
    from tests.test_program import build_program
    from yaqs_b200.params import prepare_observable_ordering
    init, instructions, obs = program_oracle_inputs(cfg)
    ntraj = cfg["ntraj"] if cfg["noise"] else 1
    with IllPosedWatch() as watch:
        want = [O.execute_program_trajectory(init, instructions, obs, cfg["seed"], t) for t in range(ntraj)]
    prog, state, noise = build_program(cfg)
    res = Y.Simulator(show_progress=False, kernels=NumpyKernels()).run(state, prog, None, noise)
    why, worst = None, 0.0
    to_sorted = prepare_observable_ordering(list(prog.observables))[1]
    if len(res.segment_results) != len(want[0][0]):
        why = f"{len(res.segment_results)} segment results, the oracle has {len(want[0][0])}"
    else:
        for k, seg in enumerate(res.segment_results):
            ref = np.stack([want[t][0][k][0] for t in range(ntraj)])           # [ntraj, n_obs (sorted), cols]
            dg = np.stack([want[t][0][k][1] for t in range(ntraj)])
            got = np.zeros_like(ref)
            for i in range(len(prog.observables)):
                rows = np.asarray(seg.trajectories[i])
                if rows.shape != got[:, to_sorted[i]].shape:
                    why = f"segment {k}: result shape {rows.shape}, the oracle has {got[:, to_sorted[i]].shape}"
                    break
                got[:, to_sorted[i]] = rows
            if why:
                break
            err = float(np.max(np.abs(got - ref))) if ref.size else 0.0
            worst = max(worst, err)
            if not err < 1e-10:
                why = f"segment {k}: observables differ by {err:.3e}"
                break
            if not (np.array_equal(seg.runtime_cost, dg[:, 0].mean(axis=0)) and np.array_equal(seg.max_bond, dg[:, 1].mean(axis=0))
                    and np.array_equal(seg.total_bond, dg[:, 2].mean(axis=0))):
                why = f"segment {k}: bond diagnostics differ"
                break
        if why is None:
            draws = np.array([want[t][2] for t in range(ntraj)])
            if not np.array_equal(np.asarray(res.stats["trajectory_draws"])[:ntraj], draws):
                why = f"stream positions differ: {res.stats['trajectory_draws']} against {draws.tolist()}"
    if why is not None and watch.notes:
        return "ill-posed: " + watch.notes[0] + " [" + why + "]", worst
    return why, worst


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=40)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--keep-going", action="store_true")
    args = ap.parse_args()
    rng = np.random.default_rng(args.seed)
    bad = 0
    for n in range(args.cases):
        cfg = draw_case(rng)
        t0 = time.time()
        try:
            why, worst = run_one(cfg)
        except Exception as exc:  # noqa: BLE001 -- a crash is a finding too
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
