"""Seeded random digital configurations (Ising Trotter circuits, BASELINE config 4's path): the batched host engine on the numpy kernel
twin against the oracle.  TEST INFRASTRUCTURE (CPU only), companion of tests/walks/fuzz_twin_vs_oracle.py.

Draws chain length, couplings, Trotter depth, periodic closure (a long-range gate), its route (gate MPO / SWAP ladder / TDVP window),
bond cap, threshold, noise (one-site and adjacent two-site processes) and sampling mode, and asks for the same jumps, identical
diagnostics, observables and pdfs <= 1e-10.

    python tests/walks/fuzz_digital_twin_vs_oracle.py --cases 50 --seed 0
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

ONE_SITE = ["lowering", "raising", "pauli_x", "pauli_y", "pauli_z"]
ADJACENT = ["crosstalk_xx", "crosstalk_yy", "crosstalk_zy", "crosstalk_xz"]


def draw_case(rng: np.random.Generator) -> dict:
    L = int(rng.integers(2, 9))
    periodic = bool(L >= 3 and rng.random() < 0.5)
    n_one = int(rng.integers(0, 3))
    noise = [(str(n), float(np.round(rng.uniform(0.01, 0.4), 3))) for n in rng.choice(ONE_SITE, size=n_one, replace=False)]
    noise2 = [(str(rng.choice(ADJACENT)), float(np.round(rng.uniform(0.01, 0.3), 3)))] if rng.random() < 0.3 else []
    return dict(L=L, J=float(np.round(rng.uniform(0.3, 1.5), 3)), g=float(np.round(rng.uniform(0.1, 1.5), 3)), dt=float(rng.choice([0.05, 0.1, 0.3])),
                timesteps=int(rng.integers(1, 5)), periodic=periodic, chi=int(rng.choice([2, 4, 8, 16])), thr=float(10.0 ** rng.uniform(-12, -5)),
                ntraj=int(rng.integers(1, 5)), seed=int(rng.integers(0, 10000)), init=str(rng.choice(["zeros", "x+", "Neel", "ones"])),
                obs=str(rng.choice(["z", "x+z", "z+zz"])), sample=bool(rng.random() < 0.6),
                gate_mode=str(rng.choice(["mpo", "swaps", "tdvp"])) if periodic else "mpo", noise=noise, noise2=noise2)


ONE_Q = ["x", "y", "z", "h", "s", "sdg", "t", "tdg", "sx", "sxdg", "id", "rx", "ry", "rz", "p", "u1", "u2", "u3", "u"]
TWO_Q = ["cx", "cz", "cp", "rxx", "ryy", "rzz", "swap"]
THREE_Q = ["ccx", "ccz", "cswap"]
N_ANGLES = {"rx": 1, "ry": 1, "rz": 1, "p": 1, "u1": 1, "u2": 2, "u3": 3, "u": 3, "cp": 1, "rxx": 1, "ryy": 1, "rzz": 1}


def draw_gate_case(rng: np.random.Generator) -> dict:
    """A random gate list over the whole registry (dag_utils.py:51-83), one gate per layer: one-, two- (adjacent and long-range, both
    site orders) and three-qubit gates, through a random gate route, from a seeded random bond-2 state (a product state makes every
    TDVP window ill-posed: min_keep = 2 keeps a numerically zero singular value)."""
    L = int(rng.integers(3, 8))
    gates = []
    for _ in range(int(rng.integers(4, 16))):
        kind = rng.random()
        if kind < 0.4:
            name = str(rng.choice(ONE_Q))
            qubits = (int(rng.integers(0, L)),)
        elif kind < 0.92 or L < 4:
            name = str(rng.choice(TWO_Q))
            a, b = (int(v) for v in rng.choice(L, size=2, replace=False))
            qubits = (a, b)
        else:
            name = str(rng.choice(THREE_Q))
            qubits = tuple(int(v) for v in rng.choice(L, size=3, replace=False))
        params = tuple(float(np.round(rng.uniform(-2.5, 2.5), 3)) for _ in range(N_ANGLES.get(name, 0)))
        gates.append((name, params, qubits))
        if rng.random() < 0.12:
            gates.append(("sample", (), ()))
    n_one = int(rng.integers(0, 3))
    noise = [(str(n), float(np.round(rng.uniform(0.01, 0.3), 3))) for n in rng.choice(ONE_SITE, size=n_one, replace=False)]
    return dict(L=L, circuit="gates", gates=gates, chi=(None if rng.random() < 0.5 else int(rng.choice([4, 8, 16]))), thr=float(10.0 ** rng.uniform(-12, -6)),
                ntraj=int(rng.integers(1, 4)), seed=int(rng.integers(0, 10000)), init=f"random:2:{int(rng.integers(1, 1000))}",
                obs=str(rng.choice(["z", "x+z", "z+zz"])), sample=bool(rng.random() < 0.6), gate_mode=str(rng.choice(["mpo", "swaps", "tdvp", "full-tdvp"])),
                noise=noise, noise2=[])


def run_one(cfg: dict):
    import torch

    import oracle.golden_cases as GC
    from oracle.np_kernels import NumpyKernels
    from tests.helpers import digital_engine_for_case
    from fuzz_twin_vs_oracle import IllPosedWatch
    with IllPosedWatch() as watch:
        traces = GC.run_oracle_digital_case(cfg)
    why, worst = _compare(cfg, traces)
    if why is not None and watch.notes:                  # a mismatch on a case whose outcome rounding decides is not a finding
        return "ill-posed: " + watch.notes[0] + " [" + why + "]", worst
    return why, worst


def _compare(cfg: dict, traces):
    import torch

    import oracle.golden_cases as GC
    from oracle.np_kernels import NumpyKernels
    from tests.helpers import digital_engine_for_case
    GC.ALL_DIGITAL_CASES["__fuzz__"] = cfg
    try:
        eng, compiled, tu, traj, _ = digital_engine_for_case("__fuzz__", NumpyKernels())
    finally:
        GC.ALL_DIGITAL_CASES.pop("__fuzz__", None)
    log = []
    results, diag = eng.run_digital(compiled.layers, tu, cfg["sample"], cfg["gate_mode"], log)
    res, dg = results.cpu().numpy(), diag.cpu().numpy()
    worst = 0.0
    for t, tr in enumerate(traces):
        if not np.array_equal(dg[t], tr.diagnostics):
            return f"diagnostics differ (trajectory {t}): engine {dg[t].tolist()} oracle {tr.diagnostics.tolist()}", worst
        err = float(np.max(np.abs(res[t] - tr.results)))
        worst = max(worst, err)
        if not err < 1e-10:
            return f"observables differ by {err:.3e} (trajectory {t})", worst
        if len(log) < len(tr.jumps):
            return f"jump log shorter than the oracle's ({len(log)} < {len(tr.jumps)})", worst
        for c, rec in enumerate(tr.jumps):
            e = log[c]
            if bool(e["jumped"][t]) != rec.jumped:
                return f"jump decision differs (trajectory {t}, call {c})", worst
            if rec.jumped:
                if int(e["choice"][t]) != rec.choice:
                    return f"jump choice differs (trajectory {t}, call {c})", worst
                pe = float(np.max(np.abs(e["pdf"][t][: len(rec.probabilities)] - np.array(rec.probabilities))))
                if not pe < 1e-10:
                    return f"pdf differs by {pe:.3e} (trajectory {t}, call {c})", worst
    if int(eng.status.abs().sum()):
        return f"status words raised: {eng.status.tolist()}", worst
    return None, worst


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=40)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--keep-going", action="store_true")
    ap.add_argument("--gates", action="store_true", help="random gate lists over the whole gate registry instead of Ising Trotter circuits")
    args = ap.parse_args()
    rng = np.random.default_rng(args.seed)
    bad = 0
    for n in range(args.cases):
        cfg = draw_gate_case(rng) if args.gates else draw_case(rng)
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
