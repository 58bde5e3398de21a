"""tests/golden/gate_library.npz: the matrix (and product-form generator, where the gate has one) of every gate name the reference
resolves through ``getattr(GateLibrary, name)`` (digital/utils/dag_utils.py:51-83, core/libraries/gate_library.py:1996-2048), written
by the UNMODIFIED reference in the build container.  TEST INFRASTRUCTURE.

    python -m oracle.make_gate_golden
"""
from __future__ import annotations

import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.ref_shim import load_reference  # noqa: E402

ANGLES = {1: [0.37], 2: [0.37, -1.21], 3: [0.37, -1.21, 2.05]}
PARAMETRIC = {"rx": 1, "ry": 1, "rz": 1, "p": 1, "u1": 1, "rxx": 1, "ryy": 1, "rzz": 1, "cp": 1, "u2": 2, "u": 3, "u3": 3}
OBSERVABLE_NAMES = ("p0", "p1", "destroy", "create", "xx", "yy", "zz")


def main() -> int:
    load_reference()
    from mqt.yaqs.core.libraries.gate_library import GateLibrary
    from mqt.yaqs.digital.utils.dag_utils import SUPPORTED_QISKIT_GATE_NAMES
    out = {"names": np.array(list(SUPPORTED_QISKIT_GATE_NAMES) + list(OBSERVABLE_NAMES))}
    for name in out["names"]:
        name = str(name)
        cls = getattr(GateLibrary, name)
        gate = cls(ANGLES[PARAMETRIC[name]]) if name in PARAMETRIC else cls()
        out[f"matrix_{name}"] = np.asarray(gate.matrix, dtype=np.complex128)
        out[f"interaction_{name}"] = np.int64(gate.interaction)
        out[f"name_{name}"] = np.array(gate.name)
        if gate.interaction >= 2:                       # the three-qubit gates assign their generator in set_sites (gate_library.py:1292-1302)
            gate.set_sites(*range(gate.interaction))
        gen = getattr(gate, "generator", None)
        if gen is not None:
            for k, f in enumerate(gen):
                out[f"generator_{name}_{k}"] = np.asarray(f, dtype=np.complex128)
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "gate_library.npz")
    np.savez_compressed(path, **out)
    print(path, len(out["names"]), "gates")
    return 0


if __name__ == "__main__":
    sys.exit(main())
