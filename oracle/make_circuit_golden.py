"""tests/golden/circuit_library.json: the gate lists the UNMODIFIED reference's circuit builders emit (core/libraries/circuit_library.py:
create_ising_circuit, create_2d_ising_circuit, create_heisenberg_circuit, create_2d_heisenberg_circuit).  qiskit is absent here, so the
module's ``QuantumCircuit`` is replaced by a recorder that notes every builder call -- name, angles, qubits, in order (barriers
included) -- which is all these functions do with it.  TEST INFRASTRUCTURE.

    python -m oracle.make_circuit_golden
"""
from __future__ import annotations

import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.ref_shim import load_reference  # noqa: E402

ANGLE_KEYS = ("theta", "phi", "lam")
QUBIT_KEYS = ("qubit", "qubit1", "qubit2", "control_qubit", "target_qubit")


class Register:
    """Stand-in for qiskit's QuantumRegister: qubit j of the k-th register of a circuit is wire offset_k + j."""

    def __init__(self, size: int, name: str = "") -> None:
        self.size, self.name, self.offset = size, name, 0

    def __getitem__(self, j: int) -> int:
        return self.offset + j


class Recorder:
    def __init__(self, *regs) -> None:
        if len(regs) == 1 and isinstance(regs[0], int):
            self.num_qubits = regs[0]
        else:
            self.num_qubits = 0
            for reg in regs:
                reg.offset = self.num_qubits
                self.num_qubits += reg.size
        self.ops: list = []

    def compose(self, other, front: bool = False, inplace: bool = True):
        assert inplace
        self.ops = (other.ops + self.ops) if front else (self.ops + other.ops)

    def barrier(self, *qubits, label=None):
        self.ops.append(["barrier", [], list(qubits), label])

    def __getattr__(self, name):
        def add(*args, **kw):
            nq = {"rx": 1, "ry": 1, "rz": 1, "p": 1, "u": 1, "rzz": 2, "rxx": 2, "ryy": 2, "cp": 2, "cx": 2, "cz": 2}[name]
            vals = list(args) + [kw[k] for k in ANGLE_KEYS if k in kw]
            qubits = [kw[k] for k in QUBIT_KEYS if k in kw]
            params = vals[: len(vals) - (nq - len(qubits))]
            qubits = vals[len(params):] + qubits
            self.ops.append([name, [float(p) for p in params], [int(q) for q in qubits], None])
        return add


CASES = {
    "ising_5": ("create_ising_circuit", dict(L=5, J=1.0, g=0.5, dt=0.1, timesteps=2)),
    "ising_4_periodic": ("create_ising_circuit", dict(L=4, J=0.7, g=1.1, dt=0.05, timesteps=1, periodic=True)),
    "ising2d_3x3": ("create_2d_ising_circuit", dict(num_rows=3, num_cols=3, J=1.0, g=0.6, dt=0.1, timesteps=2)),
    "ising2d_2x4": ("create_2d_ising_circuit", dict(num_rows=2, num_cols=4, J=0.8, g=0.3, dt=0.2, timesteps=1)),
    "heisenberg_5": ("create_heisenberg_circuit", dict(L=5, Jx=1.0, Jy=0.8, Jz=0.6, h=0.3, dt=0.1, timesteps=2)),
    "heisenberg_6_periodic": ("create_heisenberg_circuit", dict(L=6, Jx=0.5, Jy=0.4, Jz=1.2, h=0.1, dt=0.05, timesteps=1, periodic=True)),
    "heisenberg_2": ("create_heisenberg_circuit", dict(L=2, Jx=1.0, Jy=1.0, Jz=1.0, h=0.0, dt=0.1, timesteps=1)),
    "heisenberg2d_3x3": ("create_2d_heisenberg_circuit", dict(num_rows=3, num_cols=3, Jx=1.0, Jy=0.9, Jz=0.7, h=0.2, dt=0.1, timesteps=1)),
    "heisenberg2d_4x2": ("create_2d_heisenberg_circuit", dict(num_rows=4, num_cols=2, Jx=0.3, Jy=0.5, Jz=0.7, h=0.4, dt=0.1, timesteps=2)),
    "fermi_hubbard_1d_3": ("create_1d_fermi_hubbard_circuit", dict(L=3, u=0.5, t=1.0, mu=0.3, num_trotter_steps=2, dt=0.1, timesteps=1)),
    "fermi_hubbard_1d_4": ("create_1d_fermi_hubbard_circuit", dict(L=4, u=1.5, t=0.7, mu=-0.2, num_trotter_steps=1, dt=0.05, timesteps=2)),
    "fermi_hubbard_2d_2x2": ("create_2d_fermi_hubbard_circuit", dict(Lx=2, Ly=2, u=0.5, t=1.0, mu=0.3, num_trotter_steps=1, dt=0.1, timesteps=1)),
    "fermi_hubbard_2d_3x2": ("create_2d_fermi_hubbard_circuit", dict(Lx=3, Ly=2, u=1.5, t=0.7, mu=-0.2, num_trotter_steps=2, dt=0.05, timesteps=1)),
    "fermi_hubbard_2d_2x3": ("create_2d_fermi_hubbard_circuit", dict(Lx=2, Ly=3, u=1.0, t=0.4, mu=0.1, num_trotter_steps=1, dt=0.1, timesteps=2)),
    "random_nn_5": ("nearest_neighbour_random_circuit", dict(n_qubits=5, layers=4, seed=42)),
    "random_nn_6": ("nearest_neighbour_random_circuit", dict(n_qubits=6, layers=3, seed=7)),
}


def main() -> int:
    load_reference()
    import mqt.yaqs.core.libraries.circuit_library as lib
    lib.QuantumCircuit = Recorder
    lib.QuantumRegister = Register
    out = {}
    for key, (fn, kw) in CASES.items():
        circ = getattr(lib, fn)(**kw)
        out[key] = {"builder": fn, "kwargs": kw, "num_qubits": circ.num_qubits, "ops": circ.ops}
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "circuit_library.json")
    with open(path, "w") as fh:
        json.dump(out, fh)
    print(path, {k: len(v["ops"]) for k, v in out.items()})
    return 0


if __name__ == "__main__":
    sys.exit(main())
