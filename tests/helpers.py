"""Shared test helpers: build a BatchedTJM from a golden case and compare its outputs with the fixture."""
import os

import numpy as np
import torch

from oracle.golden_cases import ALL_CASES, oracle_inputs, scheduled_jumps_for
from yaqs_b200.engine import BatchedTJM, EngineParams
from yaqs_b200.rng import numpy_tables


def engine_for_case(name, kernels, trajectories=None, **kw):
    cfg = ALL_CASES[name]
    init, mpo, noise, p, obs = oracle_inputs(cfg)
    if cfg["model"] == "pauli_sum":                  # the PRODUCT's Pauli-sum MPO (exact finite-state construction), not the oracle's direct sum
        from yaqs_b200.containers import MPO
        built = MPO()
        built.from_pauli_sum(terms=cfg["terms"], length=cfg["L"])
        mpo = built.tensors
    traj = list(range(cfg["ntraj"])) if trajectories is None else list(trajectories)
    ep = EngineParams(dt=p.dt, n_steps=p.n_steps, order=p.order, max_bond_dim=p.max_bond_dim, svd_threshold=p.svd_threshold,
                      trunc_mode=p.trunc_mode, krylov_tol=p.krylov_tol, tdvp_mode=p.tdvp_mode, tdvp_sweeps=p.tdvp_sweeps,
                      sample_timesteps=p.sample_timesteps)
    eng = BatchedTJM(kernels, cfg["L"], len(traj), mpo, ep, noise, obs, init.tensors, scheduled_jumps=scheduled_jumps_for(cfg), **kw)
    tu, su = numpy_tables(cfg["seed"], traj, ep.n_times, ep.order)
    dev = kernels.device
    tu = torch.from_numpy(tu).to(dev)
    su = torch.from_numpy(su).to(dev) if su is not None else None
    return eng, tu, su, traj


def check_against_golden(name, golden_dir, results, diag, log, traj, obs_tol=1e-10, pdf_tol=1e-10):
    """Same seeds => same jump sequence, observables within obs_tol, jump pdfs within pdf_tol, identical diagnostics (bond dimensions).
    On the kernel twin the pdfs agree with the reference's to 3e-15 over all fixtures; 1e-10 is BASELINE.json's bar."""
    g = np.load(os.path.join(golden_dir, f"{name}.npz"))
    res, dg = results.cpu().numpy(), diag.cpu().numpy()
    cfg = ALL_CASES.get(name, {})
    fixture_row = {tid: r for r, tid in enumerate(cfg.get("traj_ids", range(g["results"].shape[0])))}      # global trajectory index -> fixture row
    checked = 0
    for row, tid in enumerate(traj):
        if tid not in fixture_row:                    # a trajectory of the batch the fixture does not hold
            continue
        t = fixture_row[tid]
        checked += 1
        assert np.array_equal(dg[row], g["diagnostics"][t]), (name, t, dg[row], g["diagnostics"][t])
        assert np.max(np.abs(res[row] - g["results"][t])) < obs_tol, (name, t, float(np.max(np.abs(res[row] - g["results"][t]))))
        n = int(g["ncalls"][t])
        if not log and not g["jumped"][t, :n].any():
            continue          # no stochastic process (scheduled jumps only): the reference still draws u1 per step, nothing depends on it
        assert len(log) >= n or n == 0
        for c in range(n):
            rec = log[c]
            assert bool(rec["jumped"][row]) == bool(g["jumped"][t, c]), (name, t, c)
            assert rec["u1"][row] == g["u1"][t, c]
            if g["jumped"][t, c]:
                assert int(rec["choice"][row]) == int(g["choice"][t, c]), (name, t, c)
                npr = rec["pdf"].shape[1]
                assert np.max(np.abs(rec["pdf"][row] - g["pdf"][t, c, :npr])) < pdf_tol
    return checked


def digital_engine_for_case(name, kernels, trajectories=None, **kw):
    """BatchedTJM + compiled layers + uniform table for a digital fixture (oracle/golden_cases.DIGITAL_CASES)."""
    from oracle.golden_cases import ALL_DIGITAL_CASES, digital_inputs
    from yaqs_b200.circuits import compile_circuit, create_ising_circuit
    cfg = ALL_DIGITAL_CASES[name]
    init, _layers, noise, p, obs = digital_inputs(cfg)
    L = cfg["L"]
    if cfg.get("circuit") in ("toffoli", "gates"):
        from oracle.golden_cases import gate_list_for
        from yaqs_b200.circuits import SAMPLE_LABEL, Circuit
        circ = Circuit(L)
        for gname, params, qubits in gate_list_for(cfg):
            if gname == "sample":
                circ.barrier(label=SAMPLE_LABEL)
            else:
                getattr(circ, gname)(*params, *qubits)
                circ.barrier()
        compiled = compile_circuit(circ)
    else:
        compiled = compile_circuit(create_ising_circuit(L, cfg["J"], cfg["g"], cfg["dt"], cfg["timesteps"], periodic=cfg["periodic"],
                                                        sample_every_step=cfg["sample"]))
    traj = list(range(cfg["ntraj"])) if trajectories is None else list(trajectories)
    ep = EngineParams(dt=1.0, n_steps=1, order=1, max_bond_dim=cfg["chi"], svd_threshold=cfg["thr"], krylov_tol=p.krylov_tol, digital=True)
    mpo = [np.eye(2, dtype=np.complex128).reshape(2, 2, 1, 1) for _ in range(L)]
    eng = BatchedTJM(kernels, L, len(traj), mpo, ep, noise if noise is not None else None, obs, init.tensors,
                     bond_factor=BatchedTJM.mpo_route_bond_factor(compiled.layers, cfg["gate_mode"], L), **kw)
    n2q = sum(len(l.even_two_qubit_gates) + len(l.odd_two_qubit_gates) for l in compiled.layers)
    tu, _ = numpy_tables(cfg["seed"], traj, max(n2q, 1), 1)
    return eng, compiled, torch.from_numpy(tu).to(kernels.device), traj, cfg
