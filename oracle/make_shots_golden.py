"""Generate tests/golden/shots_random6.npz: outcomes of the UNMODIFIED reference's MPS.measure_single_shot (mps.py:1282-1340).

TEST INFRASTRUCTURE, run once in the build container (``python -m oracle.make_shots_golden``).  Two seeded random MPS (L = 6,
chi = 4, centre 0); per state 48 shots drawn with a real numpy Generator whose uniforms are recorded by peeking the bit generator
state before every ``rng.choice`` (which consumes exactly one double), so a consumer can replay them from a table.
"""
from __future__ import annotations

import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import tjm_oracle as O  # noqa: E402
from oracle.ref_shim import load_reference  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
LENGTH, CHI, SEEDS, NSHOTS = 6, 4, (3, 4), 48


class _PeekRng:
    def __init__(self, gen, log):
        self._gen, self._log = gen, log

    def choice(self, n, p=None):
        state = self._gen.bit_generator.state
        self._log.append(float(self._gen.random()))
        self._gen.bit_generator.state = state
        return int(self._gen.choice(n, p=p))


def main():
    load_reference()
    from mqt.yaqs.core.data_structures.mps import MPS
    us, outs = [], []
    for seed in SEEDS:
        init = O.random_mps(LENGTH, CHI, seed)
        mps = MPS(LENGTH, state="zeros")
        mps.tensors = [t.copy() for t in init.tensors]
        mps.set_center(0)
        gen = np.random.default_rng(1000 + seed)
        u_rows, o_row = [], []
        for _ in range(NSHOTS):
            log: list[float] = []
            o_row.append(mps.measure_single_shot("Z", rng=_PeekRng(gen, log)))
            assert len(log) == LENGTH
            u_rows.append(log)
        us.append(u_rows)
        outs.append(o_row)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "shots_random6.npz"), u=np.array(us), outcomes=np.array(outs, dtype=np.int64),
                        seeds=np.array(SEEDS), length=LENGTH, chi=CHI)
    print("outcomes[0][:12] =", outs[0][:12])


if __name__ == "__main__":
    main()
