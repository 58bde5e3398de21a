"""Every gate name the reference resolves through ``getattr(GateLibrary, name)`` (dag_utils.py:51-83, gate_library.py:1996-2048): same
matrix, interaction size, name and product-form generator as the unmodified reference wrote into tests/golden/gate_library.npz
(oracle/make_gate_golden.py); and the qiskit-free ``Circuit`` builder accepts every one of them."""
import os

import numpy as np
import pytest

from yaqs_b200.circuits import Circuit, compile_circuit
from yaqs_b200.gates import GateLibrary

ANGLES = {1: [0.37], 2: [0.37, -1.21], 3: [0.37, -1.21, 2.05]}
PARAMETRIC = {"rx": 1, "ry": 1, "rz": 1, "p": 1, "u1": 1, "rxx": 1, "ryy": 1, "rzz": 1, "cp": 1, "u2": 2, "u": 3, "u3": 3}


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "gate_library.npz"))


def _make(name):
    factory = getattr(GateLibrary, name)
    return factory(*ANGLES[PARAMETRIC[name]]) if name in PARAMETRIC else factory()


def test_every_reference_gate_name_resolves_to_the_same_matrix(golden):
    for name in map(str, golden["names"]):
        gate = _make(name)
        want = golden[f"matrix_{name}"]
        assert gate.matrix.shape == want.shape, name
        assert np.max(np.abs(gate.matrix - want)) < 1e-15, name
        assert gate.interaction == int(golden[f"interaction_{name}"]), name
        assert gate.name == str(golden[f"name_{name}"]), name
        k = 0
        while f"generator_{name}_{k}" in golden:
            assert gate.generator is not None, name
            assert np.max(np.abs(np.asarray(gate.generator[k]) - golden[f"generator_{name}_{k}"])) < 1e-15, (name, k)
            k += 1
        if k == 0:
            assert getattr(gate, "generator", None) is None, name


def test_circuit_builder_takes_every_supported_gate(golden):
    width = {1: 1, 2: 2, 3: 3}
    circ = Circuit(3)
    names = [n for n in map(str, golden["names"]) if n not in ("p0", "p1", "destroy", "create", "xx", "yy", "zz")]
    for name in names:
        nq = width[int(golden[f"interaction_{name}"])]
        getattr(circ, name)(*ANGLES[PARAMETRIC[name]], *range(nq)) if name in PARAMETRIC else getattr(circ, name)(*range(nq))
        circ.barrier()
    layers = compile_circuit(circ).layers
    gates = [g for l in layers for g in (*l.single_qubit_gates, *l.even_two_qubit_gates, *l.odd_two_qubit_gates)]
    assert len(gates) == len(names)
    for gate, name in zip(gates, names):
        assert np.max(np.abs(gate.matrix - golden[f"matrix_{name}"])) < 1e-15, name


def test_position_observable_and_custom_alias():
    """gate_library.py:1845-1870 / :2042-2048: ``Observable("position", site, positions=...)`` is the diagonal operator of the supplied
    basis; ``GateLibrary.custom`` is the base class."""
    from yaqs_b200 import Observable
    from yaqs_b200.gates import BaseGate
    obs = Observable("position", 2, positions=[-0.5, 0.5])
    assert obs.gate.name == "position" and obs.sites == 2 and np.array_equal(obs.gate.matrix, np.diag([-0.5, 0.5]))
    assert GateLibrary.custom is BaseGate
    for bad, msg in (([1j, 0.0], "only real"), ([[0.0, 1.0]], "one-dimensional"), ([np.inf, 0.0], "finite")):
        with pytest.raises(ValueError, match=msg):
            GateLibrary.position(positions=bad)
