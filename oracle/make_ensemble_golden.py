"""Generate tests/golden/ensemble_*.npz: ensemble-level references for the 3-sigma parity bar (BASELINE.json north_star).

TEST INFRASTRUCTURE, run once in the build container (``python -m oracle.make_ensemble_golden``).  For every case the
UNMODIFIED reference provides
  * ``exact``   -- the Lindblad master-equation solution (analog/lindblad.py: preprocess_lindblad + lindblad_evolve),
  * ``ref_mean`` / ``ref_sem`` -- mean and standard error of its own TJM ensemble (analog_tjm_2 with its seeded numpy streams).
The tests compare OUR ensemble, driven by the device Philox generator (statistically independent of the reference's
streams), with both.
"""
from __future__ import annotations

import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle.ref_shim import load_reference  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

ENSEMBLE_CASES = {
    # BASELINE config 1 in miniature (exact solution affordable: 4^6 = 4096-dimensional Liouville space)
    "ensemble_ising6": dict(model="ising", L=6, J=1.0, g=0.5, init="zeros", chi=8, thr=1e-6, order=2, dt=0.1, steps=10,
                            noise=[("lowering", 0.1), ("pauli_z", 0.1)], ref_ntraj=400, ref_seed=2024),
    # Heisenberg / amplitude damping from the Neel state (config 2 in miniature)
    "ensemble_heis6": dict(model="heisenberg", L=6, J=1.0, g=0.0, init="Neel", chi=8, thr=1e-6, order=2, dt=0.1, steps=8,
                           noise=[("lowering", 0.2)], ref_ntraj=400, ref_seed=77),
}


def run_case(cfg: dict) -> dict:
    load_reference()
    import mqt.yaqs.analog.analog_tjm as ref_tjm
    from mqt.yaqs.analog.lindblad import lindblad_evolve, preprocess_lindblad
    from mqt.yaqs.core.data_structures.mpo import MPO
    from mqt.yaqs.core.data_structures.mps import MPS
    from mqt.yaqs.core.data_structures.noise_model import NoiseModel
    from mqt.yaqs.core.data_structures.simulation_parameters import AnalogSimParams, Observable
    from mqt.yaqs.core.libraries.gate_library import Z

    L = cfg["L"]
    H = MPO.ising(L, cfg["J"], cfg["g"]) if cfg["model"] == "ising" else MPO.heisenberg(L, cfg["J"], cfg["J"], cfg["J"])
    obs = [Observable(Z(), s) for s in range(L)]
    noise = NoiseModel([{"name": n, "sites": [i], "strength": g} for i in range(L) for (n, g) in cfg["noise"]])
    sim = AnalogSimParams(observables=obs, elapsed_time=cfg["dt"] * cfg["steps"], dt=cfg["dt"], num_traj=cfg["ref_ntraj"],
                          max_bond_dim=cfg["chi"], svd_threshold=cfg["thr"], order=cfg["order"], sample_timesteps=True,
                          random_seed=cfg["ref_seed"])
    # exact: Lindblad evolution of the density matrix
    psi = MPS(L, state=cfg["init"]).to_vec()
    rho = np.outer(psi, psi.conj())
    ctx = preprocess_lindblad(rho_initial=rho, h_sparse=H.to_sparse_matrix(), noise_model=noise, sim_params=sim, num_sites=L)
    exact, _, _ = lindblad_evolve(ctx)
    # the reference's own TJM ensemble
    rows = []
    for tr in range(cfg["ref_ntraj"]):
        res, _, _ = ref_tjm.analog_tjm_2((tr, MPS(L, state=cfg["init"]), noise, sim, H))
        rows.append(np.asarray(res, dtype=np.float64))
    rows = np.array(rows)
    return dict(exact=np.asarray(exact, dtype=np.float64), ref_mean=rows.mean(axis=0), ref_sem=rows.std(axis=0, ddof=1) / np.sqrt(len(rows)),
                ref_ntraj=np.int32(cfg["ref_ntraj"]), times=np.asarray(sim.times))


def main():
    for name, cfg in ENSEMBLE_CASES.items():
        out = run_case(cfg)
        dev = np.abs(out["ref_mean"] - out["exact"])
        print(f"{name}: max |ref TJM mean - Lindblad| = {dev.max():.4f}, max z = {(dev / np.maximum(out['ref_sem'], 1e-12)).max():.2f}")
        np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **out)


if __name__ == "__main__":
    main()
