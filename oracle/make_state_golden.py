"""tests/golden/state_presets.npz: the dense vectors of the UNMODIFIED reference's MPS presets (mps.py:71-301), zero-padded variants
(:409-452) and Haar-random chains drawn from a seeded generator put in place of ``np.random.default_rng`` (the reference's own tests
patch it the same way, tests/core/methods/tdvp/conftest.py:195-214).  TEST INFRASTRUCTURE.

    python -m oracle.make_state_golden
"""
from __future__ import annotations

import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.ref_shim import load_reference  # noqa: E402

PRESETS = [("zeros", {}), ("ones", {}), ("x+", {}), ("x-", {}), ("y+", {}), ("y-", {}), ("Neel", {}), ("wall", {}), ("basis", {"basis_string": "011010"})]
PADDED = [("zeros", 4), ("x+", 3), ("y-", 8), ("Neel", 2), ("wall", 4)]
HAAR = [(6, None, 11), (6, 4, 12), (9, 3, 13), (10, 16, 14), (3, 2, 15), (1, 4, 16)]


def main() -> int:
    load_reference()
    from mqt.yaqs.core.data_structures.mps import MPS
    out = {}
    for name, kw in PRESETS:
        out[f"preset_{name}"] = MPS(6, state=name, **kw).to_vec()
    for name, pad in PADDED:
        m = MPS(6, state=name, pad=pad)
        out[f"padded_{name}_{pad}"] = m.to_vec()
        out[f"padded_{name}_{pad}_bonds"] = np.array([t.shape[2] for t in m.tensors])
    orig = np.random.default_rng
    for length, pad, seed in HAAR:
        np.random.default_rng = lambda *a, _s=seed, **k: orig(_s)
        try:
            m = MPS(length, state="haar-random", pad=pad)
        finally:
            np.random.default_rng = orig
        out[f"haar_{length}_{pad}_{seed}"] = m.to_vec()
        out[f"haar_{length}_{pad}_{seed}_bonds"] = np.array([t.shape[2] for t in m.tensors])
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "state_presets.npz")
    np.savez_compressed(path, **out)
    print(path, len(out), "arrays")
    return 0


if __name__ == "__main__":
    sys.exit(main())
