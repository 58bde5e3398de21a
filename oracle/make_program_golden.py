"""Generate tests/golden/prog_*.npz: the UNMODIFIED reference walking SimulationProgram trajectories.

TEST INFRASTRUCTURE, run once in the build container (``python -m oracle.make_program_golden [names]``).  The programs of
``oracle/golden_cases.PROGRAM_CASES`` are compiled with the reference's own ``_compile_analog_segment`` (digital instructions are
assembled by hand around hand-compiled layers because qiskit is absent, SURVEY.md 8(c) shim 3), their noise models sampled with
``_sample_program_noise_models`` and every trajectory executed by the reference's ``_execute_program_trajectory``
(simulator.py:588-718).  Stored per segment k: ``res_k`` [ntraj, n_sorted_obs, cols], ``diag_k`` [ntraj, 3, cols]; the number of
uniforms each trajectory's Generator consumed (``draws``) pins the shared-stream bookkeeping."""
from __future__ import annotations

import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle.golden_cases import PROGRAM_CASES  # noqa: E402
from oracle.make_golden import reference_ising_compiled  # noqa: E402
from oracle.ref_shim import load_reference  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def run_program_case(name: str, cfg: dict) -> dict:
    load_reference()
    import mqt.yaqs.digital.digital_tjm as ref_dig
    import mqt.yaqs.simulator as ref_sim
    from mqt.yaqs.core.data_structures import simulation_program as SP
    from mqt.yaqs.core.data_structures.hamiltonian import Hamiltonian
    from mqt.yaqs.core.data_structures.mps import MPS
    from mqt.yaqs.core.data_structures.noise_model import NoiseModel
    from mqt.yaqs.core.data_structures.simulation_parameters import AnalogSimParams, DigitalSimParams, Observable
    from mqt.yaqs.core.libraries.gate_library import ZZ, GateLibrary, X, Z

    L = cfg["L"]
    obs = []
    for tok in cfg["obs"].split("+"):
        if tok == "z":
            obs += [Observable(Z(), s) for s in range(L)]
        elif tok == "x":
            obs += [Observable(X(), s) for s in range(L)]
        elif tok == "zz":
            obs += [Observable(ZZ(), [s, s + 1]) for s in range(L - 1)]
    procs = [{"name": n, "sites": [i], "strength": g} for i in range(L) for (n, g) in cfg["noise"]]
    noise = NoiseModel(procs) if procs else None
    sig = SP._StateSignature("mps", L, (2,) * L)
    hams: dict = {}
    instructions = []
    t_off = 0.0
    for k, seg in enumerate(cfg["segments"]):
        if seg["kind"] == "A":
            if seg["ham"] not in hams:
                hams[seg["ham"]] = Hamiltonian.ising(L, seg["J"], seg["g"])
            sp = AnalogSimParams(elapsed_time=seg["T"], dt=seg["dt"], num_traj=cfg["ntraj"], max_bond_dim=seg["chi"], svd_threshold=seg["thr"],
                                 order=seg["order"], sample_timesteps=seg["sample"])
            ins = SP._compile_analog_segment(SP._AnalogSegment(hams[seg["ham"]], sim_params=sp, noise_model=None), index=k, signature=sig,
                                             noise_model=noise, time_offset=t_off, observables=obs, num_traj=cfg["ntraj"], random_seed=cfg["seed"])
            t_off += float(sp.elapsed_time)
        else:
            sp = DigitalSimParams(num_traj=cfg["ntraj"], max_bond_dim=seg["chi"], svd_threshold=seg["thr"], sample_layers=seg["sample"])
            compiled = reference_ising_compiled(ref_dig, GateLibrary, L, seg["J"], seg["g"], seg["dt"], seg["timesteps"], False, seg["sample"])
            ex = SP._apply_program_settings(sp, observables=obs, num_traj=cfg["ntraj"], random_seed=cfg["seed"])
            if ex.sample_layers:
                ex.num_mid_measurements = compiled.num_mid_measurements
            ins = SP._CompiledDigitalInstruction(index=k, circuit=None, compiled_circuit=compiled, sim_params=sp, execution_params=ex,
                                                 noise_model=noise, time_offset=t_off)
        instructions.append(ins)
    compiled_prog = SP._CompiledProgram(instructions=tuple(instructions), state_signature=sig, observables=tuple(obs), get_state=False,
                                        num_traj=cfg["ntraj"], random_seed=cfg["seed"], default_noise_model=noise)
    compiled_prog = ref_sim._sample_program_noise_models(compiled_prog)
    _has_obs, _has_shots, stochastic = ref_sim._validate_compiled_program(compiled_prog)
    ntraj = cfg["ntraj"] if stochastic else 1
    # count the doubles each trajectory's Generator hands out (random() and the one inside choice())
    draws = []
    orig_make = ref_sim.make_trajectory_rng

    class _Counting:
        def __init__(self, gen):
            self._g, self.n = gen, 0

        def random(self, *a, **k):
            self.n += 1
            return self._g.random(*a, **k)

        def choice(self, *a, **k):
            self.n += 1
            return self._g.choice(*a, **k)

    out = {}
    per_seg_res = [[] for _ in instructions]
    per_seg_diag = [[] for _ in instructions]
    for t in range(ntraj):
        holder = {}

        def make(traj_idx, *, base_seed, _h=holder):
            _h["rng"] = _Counting(orig_make(traj_idx, base_seed=base_seed))
            return _h["rng"]

        ref_sim.make_trajectory_rng = make
        try:
            payloads, _final = ref_sim._execute_program_trajectory(t, MPS(L, state=cfg["init"]), compiled_prog, ntraj)
        finally:
            ref_sim.make_trajectory_rng = orig_make
        draws.append(holder["rng"].n)
        for k, (data, diag, _counts, _chk) in enumerate(payloads):
            per_seg_res[k].append(np.asarray(data, dtype=np.float64))
            per_seg_diag[k].append(np.asarray(diag, dtype=np.float64))
    for k in range(len(instructions)):
        out[f"res_{k}"] = np.array(per_seg_res[k])
        out[f"diag_{k}"] = np.array(per_seg_diag[k])
    out["draws"] = np.array(draws, dtype=np.int64)
    out["nseg"] = len(instructions)
    out["time_offsets"] = np.array([i.time_offset for i in instructions])
    return out


def main(argv: list[str]) -> int:
    for name in argv or list(PROGRAM_CASES):
        data = run_program_case(name, PROGRAM_CASES[name])
        np.savez_compressed(os.path.join(GOLDEN_DIR, f"{name}.npz"), **data)
        print(name, "draws", data["draws"].tolist(), [data[f"res_{k}"].shape for k in range(int(data["nseg"]))])
    return 0


if __name__ == "__main__":
    raise SystemExit(main(sys.argv[1:]))
