"""Seeded random configurations through the UNMODIFIED reference (import shim, build container only) and through the oracle.

TEST INFRASTRUCTURE.  The fixtures pin the oracle on a fixed table of cases; this walks the configuration space of
tests/walks/fuzz_twin_vs_oracle.py / fuzz_digital_twin_vs_oracle.py and asks, draw by draw, that the oracle reproduces what the reference
returns: observables <= 1e-10, identical diagnostics and bond dimensions, the same jump log.  Draws whose outcome rounding decides
(IllPosedWatch) are reported, not compared.  Needs /root/reference, so it cannot run on the GPU box; results in DESIGN.md section 2.

    python -m oracle.fuzz_reference_vs_oracle --cases 200 --seed 0 [--digital]
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np
import torch  # noqa: F401 -- imported before the shim registers its stand-in modules (torch inspects sys.modules on import)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "walks"))
for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
    os.environ.setdefault(_v, "1")


def compare(cfg: dict, digital: bool):
    from fuzz_twin_vs_oracle import IllPosedWatch
    import oracle.golden_cases as GC
    from oracle import make_golden as MG
    try:
        ref = (MG.run_digital_case if digital else MG.run_case)("fuzz", cfg)
    except ValueError as exc:                            # a run the reference refuses: the oracle has to refuse it with the same message
        try:
            (GC.run_oracle_digital_case if digital else GC.run_oracle_case)(cfg)
        except ValueError as ours:
            same = str(ours).split("(")[0].strip() == str(exc).split("(")[0].strip()
            return (None if same else f"both refuse the run, with different messages: {exc} / {ours}"), 0.0
        return f"the reference refuses the run ({exc}), the oracle completes it", 0.0
    with IllPosedWatch() as watch:
        traces = (GC.run_oracle_digital_case if digital else GC.run_oracle_case)(cfg)
    why, worst = None, 0.0
    for t, tr in enumerate(traces):
        if not np.array_equal(tr.diagnostics, ref["diagnostics"][t]):
            why = f"diagnostics differ (trajectory {t})"
            break
        err = float(np.max(np.abs(tr.results - ref["results"][t]))) if tr.results.size else 0.0
        worst = max(worst, err)
        if not err < 1e-10:
            why = f"observables differ by {err:.3e} (trajectory {t})"
            break
        if not np.array_equal(np.array(tr.bond_dims, dtype=np.int32), ref["bonds"][t]):
            why = f"bond dimensions differ (trajectory {t})"
            break
        n = int(ref["ncalls"][t])
        if len(tr.jumps) != n:
            why = f"{len(tr.jumps)} jump tests, the reference made {n} (trajectory {t})"
            break
        for c, rec in enumerate(tr.jumps):
            if rec.jumped != bool(ref["jumped"][t, c]) or rec.u1 != ref["u1"][t, c] or (rec.jumped and rec.choice != int(ref["choice"][t, c])):
                why = f"jump log differs (trajectory {t}, call {c})"
                break
            if rec.jumped:
                pe = float(np.max(np.abs(np.array(rec.probabilities) - ref["pdf"][t, c, : len(rec.probabilities)])))
                if not pe < 1e-10:
                    why = f"pdf differs by {pe:.3e} (trajectory {t}, call {c})"
                    break
        if why:
            break
    if why is not None and watch.notes:
        return "ill-posed: " + watch.notes[0] + " [" + why + "]", worst
    return why, worst


def compare_program(cfg: dict):
    """A random SimulationProgram through the reference's _execute_program_trajectory (oracle/make_program_golden.run_program_case) and
    through the oracle's executor: per segment observables <= 1e-10, identical diagnostics; the same number of doubles drawn."""
    from fuzz_twin_vs_oracle import IllPosedWatch
    from oracle import tjm_oracle as O
    from oracle.golden_cases import program_oracle_inputs
    from oracle.make_program_golden import run_program_case
    ref = run_program_case("fuzz", cfg)
    init, instructions, obs = program_oracle_inputs(cfg)
    ntraj = ref["res_0"].shape[0]
    why, worst = None, 0.0
    with IllPosedWatch() as watch:
        for t in range(ntraj):
            payloads, _final, draws = O.execute_program_trajectory(init, instructions, obs, cfg["seed"], t)
            if len(payloads) != int(ref["nseg"]):
                why = f"{len(payloads)} segment payloads, the reference has {int(ref['nseg'])}"
                break
            if draws != int(ref["draws"][t]):
                why = f"trajectory {t} drew {draws} doubles, the reference {int(ref['draws'][t])}"
                break
            for k, (res, diag) in enumerate(payloads):
                if res.shape != ref[f"res_{k}"][t].shape:
                    why = f"segment {k}: shape {res.shape} against {ref[f'res_{k}'][t].shape}"
                    break
                err = float(np.max(np.abs(res - ref[f"res_{k}"][t]))) if res.size else 0.0
                worst = max(worst, err)
                if not err < 1e-10:
                    why = f"segment {k}, trajectory {t}: observables differ by {err:.3e}"
                    break
                if not np.array_equal(diag, ref[f"diag_{k}"][t]):
                    why = f"segment {k}, trajectory {t}: diagnostics differ"
                    break
            if why:
                break
    if why is not None and watch.notes:
        return "ill-posed: " + watch.notes[0] + " [" + why + "]", worst
    return why, worst


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=100)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--digital", action="store_true")
    ap.add_argument("--gates", action="store_true", help="digital: random gate lists over the whole gate registry")
    ap.add_argument("--program", action="store_true", help="random SimulationPrograms through the reference's trajectory executor")
    ap.add_argument("--keep-going", action="store_true")
    args = ap.parse_args()
    args.digital = args.digital or args.gates
    if args.program:
        from fuzz_program_twin_vs_oracle import draw_case
    elif args.gates:
        from fuzz_digital_twin_vs_oracle import draw_gate_case as draw_case
    elif args.digital:
        from fuzz_digital_twin_vs_oracle import draw_case
    else:
        from fuzz_twin_vs_oracle import draw_case
    rng = np.random.default_rng(args.seed)
    bad = 0
    for n in range(args.cases):
        cfg = draw_case(rng)
        if not args.digital and not args.program and "y" in cfg["obs"].split("+"):
            cfg["obs"] = "x+z"                          # make_golden builds X, Z, ZZ, entropy and PVM observables
        t0 = time.time()
        try:
            why, worst = compare_program(cfg) if args.program else compare(cfg, args.digital)
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
