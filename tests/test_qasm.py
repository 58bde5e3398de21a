"""The qiskit-free OpenQASM 2 reader (yaqs_b200/qasm.py): the entry point behaves like the reference's ``load_circuit``
(digital/utils/qasm_utils.py:133-162) and a program runs to the same result as the circuit built gate by gate."""
import math

import numpy as np
import pytest

import yaqs_b200 as Y
from yaqs_b200.qasm import load_circuit, parse_openqasm2, qasm_version

SOURCE = """
// leading comment
OPENQASM 2.0;
include "qelib1.inc";
qreg q[3]; qreg r[2];
creg c[5];
h q;            /* whole-register broadcast */
rx(pi/2) q[1];
u3(0.1, -pi/4, 2*pi/3) r[0];
cx q[0], r[1];
cu1(0.5) q[2], q[0];
rzz(-2*0.1*1.0) q[1],q[2];
u2(sin(pi/6), sqrt(2)^2) r[1];
barrier q, r;
measure q -> c;
"""


def test_statements_become_circuit_ops():
    circ = load_circuit(SOURCE)
    got = [(op.name, op.qubits, tuple(round(p, 12) for p in op.params)) for op in circ.ops]
    assert circ.num_qubits == 5
    assert got == [("h", (0,), ()), ("h", (1,), ()), ("h", (2,), ()), ("rx", (1,), (round(math.pi / 2, 12),)),
                   ("u3", (3,), (0.1, round(-math.pi / 4, 12), round(2 * math.pi / 3, 12))), ("cx", (0, 4), ()), ("cp", (2, 0), (0.5,)),
                   ("rzz", (1, 2), (-0.2,)), ("u2", (4,), (0.5, 2.0)), ("barrier", (0, 1, 2, 3, 4), ()),
                   ("measure", (0,), ()), ("measure", (1,), ()), ("measure", (2,), ())]


def test_entry_point_follows_load_circuit(tmp_path):
    circ = Y.create_ising_circuit(3, 1.0, 0.5, 0.1, 1)
    assert load_circuit(circ) is circ                                            # a circuit object passes through
    path = tmp_path / "prog.qasm"
    path.write_text(SOURCE)
    assert len(load_circuit(path).ops) == len(load_circuit(str(path)).ops) == 13          # Path and path string
    assert qasm_version(SOURCE) == "2" and qasm_version("/* x */ OPENQASM 3;\nqubit q;") == "3" and qasm_version("qreg q[1];") is None
    with pytest.raises(ValueError, match="Expected a QuantumCircuit, a filesystem path to an OpenQASM file"):
        load_circuit("no such file and no header")
    with pytest.raises(ImportError, match="OpenQASM 3"):
        load_circuit("OPENQASM 3.0;\nqubit[2] q;")
    for bad, msg in (("OPENQASM 2.0; qreg q[2]; ch q[0],q[1];", "unsupported gate 'ch'"),
                     ("OPENQASM 2.0; qreg q[2]; reset q[0];", "'reset' is not supported"),
                     ("OPENQASM 2.0; qreg q[2]; creg c[2]; if(c==1) x q[0];", "'if' is not supported"),
                     ("OPENQASM 2.0; qreg q[2]; gate foo a { x a; } foo q[0];", "'gate' is not supported"),
                     ("OPENQASM 2.0; qreg q[2]; rx q[0];", "takes 1 parameter"),
                     ("OPENQASM 2.0; qreg q[2]; cx q[0], q[0];", "duplicate qubit"),
                     ("OPENQASM 2.0; qreg q[2]; x q[5];", "index out of range"),
                     ("OPENQASM 2.0; qreg q[2]; x p[0];", "unknown quantum argument"),
                     ("OPENQASM 2.0; qreg q[2]; rx(__import__('os')) q[0];", "unsupported parameter expression"),
                     ("OPENQASM 2.0; creg c[2];", "declares no qreg")):
        with pytest.raises(ValueError, match=msg):
            parse_openqasm2(bad)


def test_a_qasm_program_runs_like_the_circuit_it_describes(tmp_path):
    from oracle.np_kernels import NumpyKernels
    L, J, g, dt = 4, 1.0, 0.7, 0.2
    lines = ["OPENQASM 2.0;", 'include "qelib1.inc";', f"qreg q[{L}];"]
    for _step in range(2):                                                       # create_ising_circuit, written out
        lines += [f"rx({-2 * dt * g}) q[{s}];" for s in range(L)]
        lines += [f"rzz(-2*{dt}*{J}) q[{a}], q[{b}];" for a, b in ((0, 1), (2, 3), (1, 2))]
    text = "\n".join(lines)
    obs = [Y.Observable(Y.Z(), s) for s in range(L)]
    noise = Y.NoiseModel([{"name": "pauli_x", "sites": [i], "strength": 0.05} for i in range(L)])

    def run(operator):
        sp = Y.DigitalSimParams(observables=obs, num_traj=4, max_bond_dim=8, random_seed=9)
        return Y.Simulator(show_progress=False, kernels=NumpyKernels()).run(Y.State(L), operator, sp, noise)
    want = run(Y.create_ising_circuit(L, J, g, dt, 2))
    path = tmp_path / "ising.qasm"
    path.write_text(text)
    for operator in (text, path):
        got = run(operator)
        for a, b in zip(want.expectation_values, got.expectation_values):
            assert np.max(np.abs(np.asarray(a) - np.asarray(b))) < 1e-12
    prog = Y.SimulationProgram([(text, Y.DigitalSimParams(max_bond_dim=8))], observables=obs, num_traj=4, random_seed=9)
    res = Y.Simulator(show_progress=False, kernels=NumpyKernels()).run(Y.State(L), prog, None, noise)
    for a, b in zip(want.expectation_values, res.expectation_values):
        assert np.max(np.abs(np.asarray(a)[-1] - np.asarray(b)[-1])) < 1e-12
