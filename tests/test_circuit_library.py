"""The Trotter-circuit builders (yaqs_b200/circuits.py) emit, gate by gate and barrier by barrier, what the unmodified reference's
builders emit (core/libraries/circuit_library.py:28-349; tests/golden/circuit_library.json recorded by oracle/make_circuit_golden.py)."""
import json
import os

import numpy as np
import pytest

from yaqs_b200 import circuits

with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "circuit_library.json")) as _fh:
    GOLDEN = json.load(_fh)


@pytest.mark.parametrize("key", sorted(GOLDEN))
def test_builder_emits_the_reference_gate_list(key):
    case = GOLDEN[key]
    circ = getattr(circuits, case["builder"])(**case["kwargs"])
    assert circ.num_qubits == case["num_qubits"]
    assert len(circ.ops) == len(case["ops"]), (len(circ.ops), len(case["ops"]))
    for k, (op, (name, params, qubits, label)) in enumerate(zip(circ.ops, case["ops"])):
        assert op.name == name, (k, op, name)
        if name == "barrier":
            continue                                 # (a full barrier either way; ours lists every wire, the recorder saw none named)
        assert list(op.qubits) == qubits, (k, op, qubits)
        assert np.allclose(op.params, params, rtol=0, atol=1e-13), (k, op, params)      # (the random circuit's angles go through an expm there, a closed form here)


def test_snake_keeps_rows_contiguous_and_compiles():
    circ = circuits.create_2d_ising_circuit(3, 4, 1.0, 0.5, 0.1, 1)
    compiled = circuits.compile_circuit(circ)
    two = [g for l in compiled.layers for g in (*l.even_two_qubit_gates, *l.odd_two_qubit_gates)]
    assert len(two) == 3 * 3 + 4 * 2                  # 9 horizontal + 8 vertical bonds
    assert max(abs(g.sites[0] - g.sites[1]) for g in two) == 2 * 4 - 1
    assert sum(abs(g.sites[0] - g.sites[1]) == 1 for g in two) >= 9
