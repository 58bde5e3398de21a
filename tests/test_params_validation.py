"""AnalogSimParams / DigitalSimParams against the unmodified reference's constructors on 700 seeded random argument sets
(tests/golden/params_validation.json, oracle/make_params_golden.py): the same exception type and message -- so also the same ORDER of
checks -- or the same resolved fields (presets applied) and the same time grid (error conventions)."""
import json
import os

import numpy as np
import pytest

import yaqs_b200 as Y

with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "params_validation.json")) as _fh:
    GOLDEN = json.load(_fh)


def test_constructors_follow_the_reference():
    seen_errors = set()
    for rec in GOLDEN:
        cls = Y.AnalogSimParams if rec["cls"] == "analog" else Y.DigitalSimParams
        try:
            obj = cls(observables=[Y.Observable(Y.Z(), 0)], **rec["kwargs"])
        except NotImplementedError:
            assert rec["kwargs"].get("evolution_mode") == "bug"                  # (the one documented scope rejection)
            continue
        except Exception as exc:  # noqa: BLE001
            assert "error" in rec, (rec["kwargs"], repr(exc))
            assert [type(exc).__name__, str(exc)] == rec["error"], (rec["kwargs"], repr(exc), rec["error"])
            seen_errors.add(rec["error"][1].split(",")[0][:30])
            continue
        assert "fields" in rec, (rec["kwargs"], rec.get("error"))
        for name, want in rec["fields"].items():
            got = getattr(obj, name)
            assert got == want or (isinstance(want, float) and np.isnan(want) and np.isnan(got)), (rec["kwargs"], name, got, want)
        if "times" in rec:
            assert np.array_equal(obj.times, np.array(rec["times"])), rec["kwargs"]
    assert len(seen_errors) >= 12                      # the table really exercises the rejection paths


def test_a_cap_below_one_is_refused_when_the_run_is_set_up():
    """The reference stores any int as max_bond_dim and fails inside the first split; here the constructor stores it too and the run
    refuses it where it builds its engine parameters."""
    from oracle.np_kernels import NumpyKernels
    sp = Y.AnalogSimParams(observables=[Y.Observable(Y.Z(), 0)], elapsed_time=0.1, dt=0.1, max_bond_dim=0)
    assert sp.max_bond_dim == 0
    with pytest.raises(ValueError, match="max_bond_dim must be a positive int or None, got 0"):
        Y.Simulator(show_progress=False, kernels=NumpyKernels()).run(Y.State(3), Y.Hamiltonian.ising(3, 1.0, 0.5), sp)
