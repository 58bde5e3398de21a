"""MPS / State presets against the unmodified reference's (tests/golden/state_presets.npz, oracle/make_state_golden.py): product presets
entry by entry; zero-padded variants up to the global sign the reference's own normalize() picks up from LAPACK's QR convention when
it drops the last 1 x 1 factor (mps.py:815-839); Haar-random chains entry by entry when both draw from the same seeded generator."""
import os

import numpy as np
import pytest

from oracle.make_state_golden import HAAR, PADDED, PRESETS
from yaqs_b200.containers import MPS, State


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "state_presets.npz"))


def test_product_presets(golden):
    for name, kw in PRESETS:
        mps = MPS(6, state=name, **kw)
        assert np.max(np.abs(mps.to_vec() - golden[f"preset_{name}"])) < 1e-15, name
        assert mps.orthogonality_center == 0
        assert np.max(np.abs(State(6, initial=name, **kw).mps.to_vec() - golden[f"preset_{name}"])) < 1e-15, name


def test_padded_presets(golden):
    for name, pad in PADDED:
        mps = MPS(6, state=name, pad=pad)
        assert [t.shape[2] for t in mps.tensors] == list(golden[f"padded_{name}_{pad}_bonds"]), (name, pad)
        overlap = np.vdot(golden[f"padded_{name}_{pad}"], mps.to_vec())
        assert abs(abs(overlap) - 1.0) < 1e-14 and abs(overlap.imag) < 1e-14, (name, pad, overlap)      # the same ray; the sign is LAPACK's


def test_haar_random_draws_the_reference_state(golden, monkeypatch):
    make = np.random.default_rng
    for length, pad, seed in HAAR:
        monkeypatch.setattr(np.random, "default_rng", lambda *a, _s=seed, **k: make(_s))
        mps = MPS(length, state="haar-random", pad=pad)
        monkeypatch.setattr(np.random, "default_rng", make)
        assert [t.shape[2] for t in mps.tensors] == list(golden[f"haar_{length}_{pad}_{seed}_bonds"])
        assert np.max(np.abs(mps.to_vec() - golden[f"haar_{length}_{pad}_{seed}"])) < 1e-14, (length, pad)
        assert mps.orthogonality_center == 0 and abs(np.linalg.norm(mps.to_vec()) - 1.0) < 1e-14
        for t in mps.tensors[1:]:                                                                        # right-canonical behind the centre
            m = t.transpose(1, 0, 2).reshape(t.shape[1], -1)
            assert np.max(np.abs(m @ m.conj().T - np.eye(t.shape[1]))) < 1e-13
