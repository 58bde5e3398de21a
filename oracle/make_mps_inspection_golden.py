"""Generate tests/golden/mps_inspection.npz: what the UNMODIFIED reference's MPS inspection methods return for two seeded states.

TEST INFRASTRUCTURE.  Run here (``python -m oracle.make_mps_inspection_golden``): /root/reference does not exist on the GPU box.
The fixture holds the tensors (right-canonical, centre 0, one of them with norm 0.8) and, from the reference's own ``MPS``
(core/data_structures/mps.py): to_vec, norm, scalar_product with a second state (whole chain, one site, one pair), expect for
one-site / adjacent two-site observables (known and unknown gauge), local_expect at the centre, get_entropy / get_schmidt_spectrum of every cut,
get_max_bond / get_total_bond / get_cost / bond_dimensions, project_onto_bitstring.
"""
from __future__ import annotations

import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import tjm_oracle as O  # noqa: E402
from oracle.ref_shim import load_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "mps_inspection.npz")
L = 6
BITSTRINGS = ["000000", "010101", "110010", "111111"]


def main() -> int:
    load_reference()
    from mqt.yaqs.core.data_structures.mps import MPS
    from mqt.yaqs.core.data_structures.simulation_parameters import Observable
    from mqt.yaqs.core.libraries.gate_library import XX, YY, ZZ, X, Y, Z

    a_t = [t.copy() for t in O.random_mps(L, 4, 314).tensors]
    b_t = [t.copy() for t in O.random_mps(L, 4, 2718).tensors]       # same bonds as a: the local overlaps contract tensor with tensor
    a_t[0] = 0.8 * a_t[0]                                    # an unnormalised state: expect / norm are not divided by the norm

    def ref_mps(tensors):
        m = MPS(L, state="zeros")
        m.tensors = [t.copy() for t in tensors]
        m.set_center(0)
        return m

    a, b = ref_mps(a_t), ref_mps(b_t)
    data = {f"a{i}": t for i, t in enumerate(a_t)}
    data.update({f"b{i}": t for i, t in enumerate(b_t)})
    data["to_vec"] = a.to_vec()
    data["norm"] = np.array([a.norm(), a.norm(0), a.norm(3)])
    data["overlap"] = np.array([a.scalar_product(b), a.scalar_product(b, 2), a.scalar_product(b, [2]), a.scalar_product(b, [1, 2])])
    one = [("x", X), ("y", Y), ("z", Z)]
    data["expect_one"] = np.array([[a.expect(Observable(g(), s)) for s in range(L)] for _, g in one])
    two = [("xx", XX), ("yy", YY), ("zz", ZZ)]
    data["expect_adjacent"] = np.array([[a.expect(Observable(g(), [s, s + 1])) for s in range(L - 1)] for _, g in two])
    unknown = ref_mps(a_t)
    unknown.set_center(None)                                 # unknown gauge: the reference takes mixed_expectation (full contraction)
    data["expect_unknown_gauge"] = np.array([[unknown.expect(Observable(g(), [s, s + 1])) for s in range(L - 1)] for _, g in two])
    data["local_expect"] = np.array([a.local_expect(Observable(Z(), 0), 0), a.local_expect(Observable(XX(), [0, 1]), [0, 1])])
    data["entropy"] = np.array([a.get_entropy([s, s + 1]) for s in range(L - 1)])
    data["schmidt"] = np.array([a.get_schmidt_spectrum([s, s + 1]) for s in range(L - 1)])
    data["bonds"] = np.array([a.get_max_bond(), a.get_total_bond(), a.get_cost()] + a.bond_dimensions())
    data["pvm"] = np.array([a.project_onto_bitstring(bits) for bits in BITSTRINGS])
    np.savez_compressed(OUT, **data)
    print(f"wrote {OUT} ({os.path.getsize(OUT)} B)")
    return 0


if __name__ == "__main__":
    sys.exit(main())
