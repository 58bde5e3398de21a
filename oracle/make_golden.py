"""Generate tests/golden/*.npz by running the UNMODIFIED reference in the build container.

TEST INFRASTRUCTURE.  Run once here (``python -m oracle.make_golden``); the produced fixtures are
committed and are what travels to the GPU box (``/root/reference`` does not exist there).

For every case we record, per trajectory: observables (n_obs, T), diagnostics (3, T), the bond
dimensions of the measured state at every recorded column, and the full jump log -- one row per
``stochastic_process`` call: (dp is implied by u1/jumped; we store u1, jumped, u2, choice) plus the
normalised jump PDF.  The uniforms come from the reference's own seeded numpy streams
(``core/random_utils.py``), so a consumer can either replay the seeds or inject the recorded
uniforms.
"""
from __future__ import annotations

import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle.ref_shim import load_reference  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

from oracle.golden_cases import ALL_CASES, ALL_DIGITAL_CASES, CASES, DIGITAL_CASES, initial_state, pvm_bitstring  # noqa: E402


class _RecordingRng:
    """Wraps a numpy Generator and logs every random()/choice() the reference makes."""

    def __init__(self, gen, log, tag):
        self._gen, self._log, self._tag = gen, log, tag

    def random(self):
        u = float(self._gen.random())
        self._log.append(dict(tag=self._tag, u1=u, jumped=False, u2=np.nan, choice=-1, p=[]))
        return u

    def choice(self, n, p=None):
        # numpy's Generator.choice(n, p=p) draws exactly one uniform; peek it from a state copy
        state = self._gen.bit_generator.state
        u2 = float(self._gen.random())
        self._gen.bit_generator.state = state
        idx = int(self._gen.choice(n, p=p))
        rec = self._log[-1]
        rec.update(jumped=True, u2=u2, choice=idx, p=[float(x) for x in p])
        return idx


def run_case(name: str, cfg: dict) -> dict:
    load_reference()
    import mqt.yaqs.analog.analog_tjm as ref_tjm
    from mqt.yaqs.core.data_structures.mpo import MPO
    from mqt.yaqs.core.data_structures.mps import MPS
    from mqt.yaqs.core.data_structures.noise_model import NoiseModel
    from mqt.yaqs.core.data_structures.simulation_parameters import AnalogSimParams, Observable
    from mqt.yaqs.core.libraries.gate_library import XX, YY, ZZ, X, Y, Z
    from mqt.yaqs.core.random_utils import make_sample_rng, make_trajectory_rng

    L = cfg["L"]
    if cfg["model"] == "ising":
        H = MPO.ising(L, cfg["J"], cfg["g"])
    elif cfg["model"] == "pauli_sum":
        H = MPO()
        H.from_pauli_sum(terms=[(c, spec) for c, spec in cfg["terms"]], length=L)
    else:
        H = MPO.heisenberg(L, cfg["J"], cfg["J"], cfg["J"])
    obs = []
    for tok in cfg["obs"].split("+"):
        if tok == "z":
            obs += [Observable(Z(), s) for s in range(L)]
        elif tok == "x":
            obs += [Observable(X(), s) for s in range(L)]
        elif tok == "zz":
            obs += [Observable(ZZ(), [s, s + 1]) for s in range(L - 1)]
        elif tok == "ent":
            obs += [Observable("entropy", [s, s + 1]) for s in range(L - 1)]
        elif tok.startswith("pvm"):
            obs += [Observable(pvm_bitstring(tok, L))]          # unknown names resolve to GateLibrary.pvm(bitstring)
    procs = [{"name": n, "sites": [i], "strength": g} for i in range(L) for (n, g) in cfg["noise"]]
    procs += [{"name": n, "sites": list(s), "strength": g} for (n, s, g) in cfg.get("noise2", [])]
    sched = [{"time": t, "sites": list(st), "name": n} for (t, st, n) in cfg.get("scheduled", [])]
    noise = NoiseModel(procs, scheduled_jumps=sched or None) if (procs or sched) else None
    sim = AnalogSimParams(observables=obs, elapsed_time=cfg["dt"] * cfg["steps"], dt=cfg["dt"], num_traj=cfg["ntraj"],
                          max_bond_dim=cfg["chi"], svd_threshold=cfg["thr"], order=cfg["order"],
                          sample_timesteps=cfg["sample_timesteps"], random_seed=cfg["seed"],
                          tdvp_mode=cfg.get("tdvp_mode", "2site"), trunc_mode=cfg.get("trunc_mode", "discarded_weight"))
    assert len(sim.times) == cfg["steps"] + 1, (len(sim.times), cfg["steps"])
    # the workers see observables in sorted order (simulator.py:1561-1562)
    sorted_sites = [getattr(o, "sites", -1) for o in sim.sorted_observables]          # PVM observables carry no sites
    sorted_names = [o.gate.name for o in sim.sorted_observables]

    ncol = len(sim.times) if cfg["sample_timesteps"] else 1
    out_res, out_diag, out_bonds, logs = [], [], [], []
    orig_record = MPS.record_diagnostics
    for tr in cfg.get("traj_ids", range(cfg["ntraj"])):       # global trajectory indices (fixture rows follow this order)
        log: list[dict] = []
        bonds: list[list[int]] = []

        def rec(self, diagnostics, column_index, _b=bonds):
            _b.append([t.shape[1] for t in self.tensors[1:]])
            return orig_record(self, diagnostics, column_index)

        MPS.record_diagnostics = rec
        orig_sample_rng = ref_tjm.make_sample_rng
        ref_tjm.make_sample_rng = lambda traj_idx, *, base_seed, timestep, _log=log: _RecordingRng(
            make_sample_rng(traj_idx, base_seed=base_seed, timestep=timestep), _log, f"sample{timestep}")
        try:
            if cfg["init"].startswith("random:"):
                mps = MPS(L, state="zeros")
                mps.tensors = [t.copy() for t in initial_state(cfg).tensors]
                mps.set_center(0)
            else:
                mps = MPS(L, state=cfg["init"], pad=cfg.get("pad"))
            rng = _RecordingRng(make_trajectory_rng(tr, base_seed=cfg["seed"]), log, "traj")
            backend = ref_tjm.analog_tjm_2 if cfg["order"] == 2 else ref_tjm.analog_tjm_1
            res, diag, _ = backend((tr, mps, noise, sim, H), rng=rng)
        finally:
            MPS.record_diagnostics = orig_record
            ref_tjm.make_sample_rng = orig_sample_rng
        out_res.append(np.asarray(res, dtype=np.float64))
        out_diag.append(np.asarray(diag, dtype=np.float64))
        assert len(bonds) == ncol, (len(bonds), ncol)
        out_bonds.append(bonds)
        logs.append(log)

    ncalls = max(len(lg) for lg in logs)
    nproc = max([len(r["p"]) for lg in logs for r in lg] + [1])
    u1 = np.full((cfg["ntraj"], ncalls), np.nan)
    u2 = np.full((cfg["ntraj"], ncalls), np.nan)
    jumped = np.zeros((cfg["ntraj"], ncalls), dtype=np.int8)
    choice = np.full((cfg["ntraj"], ncalls), -1, dtype=np.int32)
    pdf = np.zeros((cfg["ntraj"], ncalls, nproc))
    is_sample = np.zeros((cfg["ntraj"], ncalls), dtype=np.int8)
    ncall_per = np.array([len(lg) for lg in logs], dtype=np.int32)
    for t, lg in enumerate(logs):
        for c, r in enumerate(lg):
            u1[t, c], u2[t, c], jumped[t, c], choice[t, c] = r["u1"], r["u2"], r["jumped"], r["choice"]
            pdf[t, c, :len(r["p"])] = r["p"]
            is_sample[t, c] = r["tag"].startswith("sample")
    return dict(results=np.array(out_res), diagnostics=np.array(out_diag), bonds=np.array(out_bonds, dtype=np.int32),
                u1=u1, u2=u2, jumped=jumped, choice=choice, pdf=pdf, is_sample=is_sample, ncalls=ncall_per,
                sorted_obs_sites=np.array([s if isinstance(s, list) else [s, -1] for s in
                                           [x if isinstance(x, list) else [x, -1] for x in sorted_sites]], dtype=np.int32),
                sorted_obs_names=np.array(sorted_names), times=np.asarray(sim.times), krylov_tol=np.float64(sim.krylov_tol))


def reference_ising_compiled(ref_dig, GateLibrary, L, J, g, dt, timesteps, periodic=False, sample=False):   # noqa: N803
    """The layers the reference's _compile_circuit produces for create_ising_circuit (one gate per layer, even/odd by lowest
    qubit), built from the reference's own GateLibrary gates (qiskit is absent here, SURVEY.md 8(c) shim 3)."""
    alpha, beta = -2 * dt * g, -2 * dt * J
    layers = []

    def layer(single=(), even=(), odd=(), pts=0):
        return ref_dig._CompiledCircuitLayer(tuple(single), tuple(even), tuple(odd), pts)

    for step in range(timesteps):
        for s in range(L):
            gate = GateLibrary.rx([alpha])
            gate.set_sites(s)
            layers.append(layer(single=[gate]))
        pairs = [(2 * s, 2 * s + 1) for s in range(L // 2)] + [(2 * s - 1, 2 * s) for s in range(1, L // 2)]
        if L % 2 != 0 and L != 1:
            pairs.append((L - 2, L - 1))
        if periodic and L > 1:
            pairs.append((0, L - 1))
        for a, b in pairs:
            gate = GateLibrary.rzz([beta])
            gate.set_sites(a, b)
            layers.append(layer(even=[gate]) if min(a, b) % 2 == 0 else layer(odd=[gate]))
        if sample and step < timesteps - 1:
            layers.append(layer(pts=1))
    return ref_dig._CompiledCircuit(tuple(layers), sum(l.sample_points for l in layers))


def run_digital_case(name: str, cfg: dict) -> dict:
    """Run the reference's digital_tjm (digital/digital_tjm.py:636) on a hand-compiled Ising Trotter circuit.  qiskit is
    absent here, so the DAG front end is bypassed exactly as SURVEY.md 8(c) shim 3 describes: the layers the reference's
    _compile_circuit would produce for create_ising_circuit (one gate per layer, even/odd by lowest qubit) are built
    from the reference's own GateLibrary gates and passed through ``compiled_circuit=``."""
    load_reference()
    import mqt.yaqs.digital.digital_tjm as ref_dig
    from mqt.yaqs.core.data_structures.mps import MPS
    from mqt.yaqs.core.data_structures.noise_model import NoiseModel
    from mqt.yaqs.core.data_structures.simulation_parameters import DigitalSimParams, Observable
    from mqt.yaqs.core.libraries.gate_library import ZZ, GateLibrary, X, Z
    from mqt.yaqs.core.random_utils import make_trajectory_rng

    L = cfg["L"]
    if cfg.get("circuit") in ("toffoli", "gates"):           # gates from the reference's own GateLibrary, one gate per layer
        from oracle.golden_cases import gate_list_for
        layers = []
        for gname, params, qubits in gate_list_for(cfg):
            if gname == "sample":
                layers.append(ref_dig._CompiledCircuitLayer((), (), (), 1))
                continue
            gate = getattr(GateLibrary, gname)(list(params)) if params else getattr(GateLibrary, gname)()
            gate.set_sites(*qubits)
            if len(qubits) == 1:
                layers.append(ref_dig._CompiledCircuitLayer((gate,), (), (), 0))
            elif min(qubits) % 2 == 0:
                layers.append(ref_dig._CompiledCircuitLayer((), (gate,), (), 0))
            else:
                layers.append(ref_dig._CompiledCircuitLayer((), (), (gate,), 0))
        compiled = ref_dig._CompiledCircuit(tuple(layers), sum(l.sample_points for l in layers))
    else:
        compiled = reference_ising_compiled(ref_dig, GateLibrary, L, cfg["J"], cfg["g"], cfg["dt"], cfg["timesteps"], cfg["periodic"], cfg["sample"])
    mid = compiled.num_mid_measurements

    obs = []
    for tok in cfg["obs"].split("+"):
        if tok == "z":
            obs += [Observable(Z(), s) for s in range(L)]
        elif tok == "x":
            obs += [Observable(X(), s) for s in range(L)]
        elif tok == "zz":
            obs += [Observable(ZZ(), [s, s + 1]) for s in range(L - 1)]
    procs = [{"name": n, "sites": [i], "strength": g} for i in range(L) for (n, g) in cfg["noise"]]
    procs += [{"name": n, "sites": [i, i + 1], "strength": g} for i in range(L - 1) for (n, g) in cfg.get("noise2", [])]
    noise = NoiseModel(procs) if procs else None
    sim = DigitalSimParams(observables=obs, num_traj=cfg["ntraj"], max_bond_dim=cfg["chi"], svd_threshold=cfg["thr"],
                           sample_layers=cfg["sample"], random_seed=cfg["seed"], gate_mode=cfg["gate_mode"])
    sim.num_mid_measurements = mid
    out_res, out_diag, out_bonds, logs = [], [], [], []
    orig_record = MPS.record_diagnostics
    for tr in cfg.get("traj_ids", range(cfg["ntraj"])):       # global trajectory indices (fixture rows follow this order)
        log: list[dict] = []
        bonds: list[list[int]] = []

        def rec(self, diagnostics, column_index, _b=bonds):
            _b.append([t.shape[1] for t in self.tensors[1:]])
            return orig_record(self, diagnostics, column_index)

        MPS.record_diagnostics = rec
        try:
            if cfg["init"].startswith("random:"):
                mps = MPS(L, state="zeros")
                mps.tensors = [t.copy() for t in initial_state(cfg).tensors]
                mps.set_center(0)
            else:
                mps = MPS(L, state=cfg["init"])
            rng = _RecordingRng(make_trajectory_rng(tr, base_seed=cfg["seed"]), log, "traj")
            res, diag, _counts, _final = ref_dig.digital_tjm((tr, mps, noise, sim, None), rng=rng, compiled_circuit=compiled)
        finally:
            MPS.record_diagnostics = orig_record
        out_res.append(np.asarray(res, dtype=np.float64))
        out_diag.append(np.asarray(diag, dtype=np.float64))
        out_bonds.append(bonds)
        logs.append(log)
    ncalls = max([len(lg) for lg in logs] + [1])
    nproc = max([len(r["p"]) for lg in logs for r in lg] + [1])
    nt = cfg["ntraj"]
    u1 = np.full((nt, ncalls), np.nan)
    u2 = np.full((nt, ncalls), np.nan)
    jumped = np.zeros((nt, ncalls), dtype=np.int8)
    choice = np.full((nt, ncalls), -1, dtype=np.int32)
    pdf = np.zeros((nt, ncalls, nproc))
    for t, lg in enumerate(logs):
        for c, r in enumerate(lg):
            u1[t, c], u2[t, c], jumped[t, c], choice[t, c] = r["u1"], r["u2"], r["jumped"], r["choice"]
            pdf[t, c, :len(r["p"])] = r["p"]
    return dict(results=np.array(out_res), diagnostics=np.array(out_diag), bonds=np.array(out_bonds, dtype=np.int32), u1=u1, u2=u2,
                jumped=jumped, choice=choice, pdf=pdf, ncalls=np.array([len(lg) for lg in logs], dtype=np.int32),
                sorted_obs_names=np.array([o.gate.name for o in sim.sorted_observables]))


def main(argv: list[str]) -> int:
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    names = argv or (list(CASES) + list(DIGITAL_CASES))
    for name in names:
        data = run_digital_case(name, ALL_DIGITAL_CASES[name]) if name in ALL_DIGITAL_CASES else run_case(name, ALL_CASES[name])
        path = os.path.join(GOLDEN_DIR, f"{name}.npz")
        np.savez_compressed(path, **data)
        njump = int(data["jumped"].sum())
        print(f"{name}: results {data['results'].shape} jumps {njump}/{int(data['ncalls'].sum())} "
              f"max bond {int(data['bonds'].max()) if data['bonds'].size else 0} -> {os.path.getsize(path)} B")
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
