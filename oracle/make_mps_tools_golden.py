"""tests/golden/mps_tools.npz: what the UNMODIFIED reference's host MPS utilities return on seeded random chains -- single-shot outcomes
in the three bases drawn from seeded generators (mps.py:1282-1341), ``compress`` in the four truncation modes (:841-899: bond
dimensions, dense vector, centre), ``check_canonical_form`` after ``set_canonical_form`` (:766-813, 1581-1631).  TEST INFRASTRUCTURE.

    python -m oracle.make_mps_tools_golden
"""
from __future__ import annotations

import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import tjm_oracle as O  # noqa: E402
from oracle.ref_shim import load_reference  # noqa: E402

CHAINS = [(5, 4, 11), (7, 8, 12), (3, 2, 13), (6, 3, 14), (2, 2, 15), (8, 4, 16)]            # (length, bond, seed of oracle.random_mps)
COMPRESS = [("discarded_weight", 1e-3, None), ("relative", 5e-2, None), ("hard_cutoff", 1e-1, None), ("relative_discarded_weight", 1e-2, None),
            ("discarded_weight", 1e-12, 2), ("discarded_weight", 1e-12, 1)]


def main() -> int:
    load_reference()
    from mqt.yaqs.core.data_structures.mps import MPS
    out = {}
    for k, (length, chi, seed) in enumerate(CHAINS):
        tensors = O.random_mps(length, chi, seed).tensors
        mps = MPS(length, tensors=[t.copy() for t in tensors])
        mps.set_center(0)
        out[f"shots_{k}"] = np.array([[mps.measure_single_shot(basis, rng=np.random.default_rng(100 * k + s)) for s in range(12)] for basis in "ZXY"])
        centre = (k * 2) % length
        mps.set_canonical_form(centre)
        out[f"canonical_{k}"] = np.array(mps.check_canonical_form())
        mode, thr, cap = COMPRESS[k]
        mps.compress(thr, max_bond_dim=cap, trunc_mode=mode)
        out[f"compressed_bonds_{k}"] = np.array([t.shape[2] for t in mps.tensors])
        out[f"compressed_vec_{k}"] = mps.to_vec()
        out[f"compressed_centre_{k}"] = np.int64(-1 if mps.orthogonality_center is None else mps.orthogonality_center)
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "mps_tools.npz")
    np.savez_compressed(path, **out)
    print(path, len(out), "arrays")
    return 0


if __name__ == "__main__":
    sys.exit(main())
