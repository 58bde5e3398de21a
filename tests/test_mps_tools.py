"""Gauge moves, compression and projective sampling of a host MPS (yaqs_b200/containers.py) against the unmodified reference
(tests/golden/mps_tools.npz, oracle/make_mps_tools_golden.py) and against what they must preserve."""
import os

import numpy as np
import pytest

from oracle import tjm_oracle as O
from oracle.make_mps_tools_golden import CHAINS, COMPRESS
from yaqs_b200.containers import MPS


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "mps_tools.npz"))


def _chain(k):
    length, chi, seed = CHAINS[k]
    return MPS(length, tensors=[t.copy() for t in O.random_mps(length, chi, seed).tensors])


@pytest.mark.parametrize("k", range(len(CHAINS)))
def test_against_the_reference(golden, k):
    mps = _chain(k)
    vec = mps.to_vec().copy()
    shots = np.array([[mps.measure_single_shot(basis, rng=np.random.default_rng(100 * k + s)) for s in range(12)] for basis in "ZXY"])
    assert np.array_equal(shots, golden[f"shots_{k}"])                           # the same generator calls, the same conditional probabilities
    assert np.max(np.abs(mps.to_vec() - vec)) < 1e-13                            # sampling works on a copy
    centre = (k * 2) % mps.length
    mps.set_canonical_form(centre)
    assert mps.check_canonical_form() == list(golden[f"canonical_{k}"]) and centre in mps.check_canonical_form()
    assert np.max(np.abs(mps.to_vec() - vec)) < 1e-13
    mode, thr, cap = COMPRESS[k]
    mps.compress(thr, max_bond_dim=cap, trunc_mode=mode)
    assert [t.shape[2] for t in mps.tensors] == list(golden[f"compressed_bonds_{k}"])
    want = golden[f"compressed_vec_{k}"]
    assert abs(abs(np.vdot(want, mps.to_vec())) - np.linalg.norm(want) ** 2) < 1e-10 and abs(np.linalg.norm(mps.to_vec()) - np.linalg.norm(want)) < 1e-12
    assert mps.orthogonality_center == int(golden[f"compressed_centre_{k}"])


def test_moves_preserve_the_vector_and_track_the_centre():
    mps = _chain(1)
    vec = mps.to_vec().copy()
    for target in (3, 0, 6, 2):
        mps.shift_center_to(target)
        assert mps.orthogonality_center == target and target in mps.check_canonical_form()
        assert np.max(np.abs(mps.to_vec() - vec)) < 1e-13
    mps.flip_network()
    assert mps.orthogonality_center == mps.length - 1 - 2 and mps.flipped
    flipped = mps.to_vec().reshape((2,) * mps.length).transpose(*reversed(range(mps.length))).reshape(-1)      # qubit order reversed
    assert np.max(np.abs(flipped - vec)) < 1e-13
    mps.flip_network()
    assert np.max(np.abs(mps.to_vec() - vec)) < 1e-13 and not mps.flipped
    counts = mps.measure_shots(50, rng=np.random.default_rng(1))
    assert sum(counts.values()) == 50 and all(0 <= key < 2 ** mps.length for key in counts)
    with pytest.raises(ValueError, match="Invalid basis"):
        mps.measure_single_shot("W")
