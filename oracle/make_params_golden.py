"""tests/golden/params_validation.json: what the UNMODIFIED reference's AnalogSimParams / DigitalSimParams constructors do with seeded
random argument sets -- the exception type and message, or the resolved fields (presets applied) and the time grid.  The order in which
arguments are checked is part of the drop-in boundary's error conventions.  TEST INFRASTRUCTURE.

    python -m oracle.make_params_golden
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.ref_shim import load_reference  # noqa: E402

ANALOG_FIELDS = ("num_traj", "max_bond_dim", "svd_threshold", "krylov_tol", "order", "dt", "elapsed_time", "trunc_mode", "tdvp_mode", "tdvp_sweeps",
                 "sample_timesteps", "preset", "random_seed", "get_state")
DIGITAL_FIELDS = ("num_traj", "max_bond_dim", "svd_threshold", "krylov_tol", "trunc_mode", "tdvp_mode", "tdvp_sweeps", "gate_mode", "shots", "sample_layers",
                  "num_mid_measurements", "preset", "random_seed", "get_state")


def draw(rng, analog: bool) -> dict:
    def pick(*opts):
        return opts[int(rng.integers(0, len(opts)))]
    kw = {}
    if analog:
        if rng.random() < 0.7:
            kw["elapsed_time"] = pick(0.1, 1.0, 0.35, 0.0, -1.0, 1e-9, 3, float("nan"), float("inf"), "1", True, 0.30000000000000004, 10.0)
        if rng.random() < 0.7:
            kw["dt"] = pick(0.1, 0.05, 0.0, -0.1, 0.3, 1, float("nan"), "x", 0.1 + 1e-12, 0.25)
        if rng.random() < 0.4:
            kw["order"] = pick(1, 2)
        if rng.random() < 0.3:
            kw["sample_timesteps"] = pick(True, False)
        if rng.random() < 0.15:
            kw["evolution_mode"] = pick("tdvp", "krylov")
    else:
        if rng.random() < 0.4:
            kw["gate_mode"] = pick("mpo", "tdvp", "full-tdvp", "swaps", "tebd")
        if rng.random() < 0.3:
            kw["shots"] = pick(None, 10, 0, -1, 2.5, True)
        if rng.random() < 0.3:
            kw["sample_layers"] = pick(True, False)
        if rng.random() < 0.3:
            kw["num_mid_measurements"] = pick(0, 3)
    if rng.random() < 0.4:
        kw["num_traj"] = pick(1, 10, None)
    if rng.random() < 0.5:
        kw["max_bond_dim"] = pick(None, 1, 16, 2.5, "8")
    if rng.random() < 0.4:
        kw["trunc_mode"] = pick("discarded_weight", "relative", "hard_cutoff", "relative_discarded_weight", "foo", 3)
    if rng.random() < 0.4:
        kw["svd_threshold"] = pick(None, 1e-6, 0.0, -1e-3, float("nan"), float("inf"))
    if rng.random() < 0.4:
        kw["krylov_tol"] = pick(None, 1e-4, 0.0, -1.0, float("nan"))
    if rng.random() < 0.4:
        kw["preset"] = pick("fast", "balanced", "accurate", "exact", "turbo")
    if rng.random() < 0.3:
        kw["random_seed"] = pick(None, 0, 5, -1, 1.5, True, "7")
    if rng.random() < 0.3:
        kw["tdvp_sweeps"] = pick(1, 2, 0, -1, 1.5, True)
    if rng.random() < 0.3:
        kw["tdvp_mode"] = pick("1site", "2site", "dynamic", "3site")
    if rng.random() < 0.2:
        kw["get_state"] = pick(True, False)
    return kw


def main() -> int:
    load_reference()
    from mqt.yaqs.core.data_structures.simulation_parameters import AnalogSimParams, DigitalSimParams, Observable
    from mqt.yaqs.core.libraries.gate_library import Z
    rng = np.random.default_rng(2026)
    out = []
    for k in range(700):
        analog = k % 2 == 0
        kw = draw(rng, analog)
        cls, fields = (AnalogSimParams, ANALOG_FIELDS) if analog else (DigitalSimParams, DIGITAL_FIELDS)
        rec = {"cls": "analog" if analog else "digital", "kwargs": kw}
        try:
            obj = cls(observables=[Observable(Z(), 0)], **kw)
        except Exception as exc:  # noqa: BLE001
            rec["error"] = [type(exc).__name__, str(exc)]
        else:
            rec["fields"] = {f: getattr(obj, f) for f in fields}
            if analog:
                rec["times"] = [float(t) for t in obj.times]
        out.append(rec)
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "params_validation.json")
    with open(path, "w") as fh:
        json.dump(out, fh)
    print(path, len(out), "argument sets,", sum("error" in r for r in out), "rejected")
    return 0


if __name__ == "__main__":
    sys.exit(main())
