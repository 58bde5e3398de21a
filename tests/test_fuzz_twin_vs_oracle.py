"""A short seeded walk through the configuration space (tests/walks/fuzz_*.py): the batched host engine on the numpy kernel twin against the
oracle on configurations no fixture holds.  The long walks (hundreds of cases per path) are run by hand, results in DESIGN.md section 2;
this keeps the tools alive and adds a few dozen unseen configurations to every CPU run."""
import os
import sys

import numpy as np
import pytest

TOOLS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "walks")
if TOOLS not in sys.path:
    sys.path.insert(0, TOOLS)


def _walk(module, ncases, seed, routes=(None,)):
    rng = np.random.default_rng(seed)
    ill = 0
    for n in range(ncases):
        cfg = module.draw_case(rng)
        for route in routes:
            why, worst = module.run_one(cfg, route) if route != "n/a" else module.run_one(cfg)
            if why is not None and why.startswith("ill-posed"):
                ill += 1
                continue
            assert why is None, (n, route, cfg, why)
    assert ill <= ncases                                # (ill-posed draws are reported by the tools, they are not failures)


def test_analog_walk():
    import fuzz_twin_vs_oracle as F
    _walk(F, 10, 2026, routes=(None, 0))


def test_digital_walk():
    import fuzz_digital_twin_vs_oracle as F
    _walk(F, 15, 2027, routes=("n/a",))


def test_program_walk():
    import fuzz_program_twin_vs_oracle as F
    _walk(F, 15, 2028, routes=("n/a",))


def test_ill_posed_watch_flags_a_degenerate_cut():
    """Two equal singular values with the cap between them: whatever is kept is LAPACK's choice (found by the walk: max_bond_dim = 1
    behind an adjacent two-site jump on a left-canonical pair, singular values (sqrt 2, sqrt 2))."""
    import fuzz_twin_vs_oracle as F
    from oracle import tjm_oracle as O
    with F.IllPosedWatch() as watch:
        assert O.truncate(np.array([np.sqrt(2.0), np.sqrt(2.0)]), "discarded_weight", 1e-6, 1) == 1
        assert O.truncate(np.array([1.0, 0.5, 1e-17]), "discarded_weight", 1e-12, None, 3) == 3
        assert O.truncate(np.array([1.0, 0.5, 0.1]), "discarded_weight", 1e-6, 2) == 2
    assert len(watch.notes) == 2 and "equal singular values" in watch.notes[0] and "numerically zero" in watch.notes[1]
