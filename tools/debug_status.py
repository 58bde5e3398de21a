"""Diagnostic: run the C3 workload (small batch) through BatchedTJM.run and report every Jacobi SVD call that raises a status
bit: which engine call it came from, matrix sizes, sweeps used, and a dump of the offending matrix for offline analysis."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from bench import saturated_tensors  # noqa: E402
from yaqs_b200 import Hamiltonian, NoiseModel, State  # noqa: E402
from yaqs_b200 import engine as E  # noqa: E402
from yaqs_b200.kernels import CudaKernels  # noqa: E402
from yaqs_b200.rng import numpy_tables  # noqa: E402


def main():
    traj = [int(t) for t in os.environ.get("DBG_TRAJ", "").split(",") if t.strip()]
    B = len(traj) if traj else int(os.environ.get("DBG_BATCH", "48"))
    traj = traj or list(range(B))
    L = int(os.environ.get("DBG_L", "100"))
    chi = int(os.environ.get("DBG_CHI", "128"))
    steps = int(os.environ.get("DBG_STEPS", "2"))
    dev = torch.device("cuda:0")
    kern = CudaKernels(dev)
    ham = Hamiltonian.ising(L, J=1.0, g=0.5)
    state = State(tensors=saturated_tensors(L, chi))
    state.mps.normalize()
    noise = NoiseModel([{"name": n, "sites": [i], "strength": 0.1} for i in range(L) for n in ("lowering", "pauli_z")])
    obs = [(np.array([[1, 0], [0, -1]], dtype=np.complex128), [s]) for s in range(L)]
    ep = E.EngineParams(dt=0.1, n_steps=steps, order=2, max_bond_dim=chi, svd_threshold=1e-6, krylov_tol=1e-4)
    eng = E.BatchedTJM(kern, L, B, ham.mpo.tensors, ep, noise.processes, obs, state.mps.tensors)
    tu, su = numpy_tables(42, traj, ep.n_times, 2)
    tu, su = torch.from_numpy(tu).to(dev), torch.from_numpy(su).to(dev)
    events = []
    ctx = {"call": "?", "tdvp_calls": 0}
    orig_jac = eng._svd_jacobi

    def jac(mode, src, el, er, trunc, dist, mask, dims=None):
        before = eng.status.clone()
        orig_jac(mode, src, el, er, trunc, dist, mask, dims)
        new = (eng.status != before).nonzero().flatten().tolist()
        if new:
            meta = eng.svd_meta.view(B, -1).cpu().numpy()
            s = eng.svd_s.view(B, -1).cpu().numpy()
            for b in new[:4]:
                ev = {"where": ctx["call"], "tdvp_calls": ctx["tdvp_calls"], "global_traj": traj[b], "traj": b, "status": int(eng.status[b]), "mode": int(mode), "dist": dist, "trunc": [float(x) if isinstance(x, float) else x for x in trunc],
                      "meta": meta[b].tolist(), "s_head": s[b][:4].tolist(), "s_min": float(np.nanmin(s[b][:max(1, min(meta[b][3], meta[b][4]))])),
                      "nan_in_s": bool(np.isnan(s[b]).any()), "fallback": mask is not None}
                events.append(ev)
                print(json.dumps(ev), flush=True)
                if len(events) <= 6:
                    os.makedirs("gpurun_out", exist_ok=True)
                    eng.K.svd_prepare(B, mask, mode, el, er, eng.d, src, eng._buf(eng.svd_w, eng.lay.svd_dim ** 2), eng._buf(eng.svd_q, eng.lay.svd_dim ** 2),
                                      eng.svd_meta, eng.svd_rot)          # the matrix as it entered the Jacobi sweeps
                    k_, len_ = int(meta[b][0]), int(meta[b][1])
                    np.save(f"gpurun_out/dbg_w_{len(events)}.npy", eng.svd_w.view(B, -1)[b][: k_ * len_].cpu().numpy().reshape(k_, len_))
    eng._svd_jacobi = jac
    for name in ("tdvp", "apply_dissipation", "stochastic_process", "_apply_jumps", "evaluate_observables"):
        f = getattr(eng, name)

        def wrapped(*a, _f=f, _n=name, **k):
            prev = ctx["call"]
            if _n == "tdvp":
                ctx["tdvp_calls"] += 1
            ctx["call"] = _n if prev == "?" else prev + ">" + _n
            try:
                return _f(*a, **k)
            finally:
                ctx["call"] = prev
        setattr(eng, name, wrapped)
    r, dg = eng.run(tu, su, None)
    torch.cuda.synchronize()
    st = eng.status.cpu().numpy()
    print(json.dumps({"state_nan": bool(torch.isnan(torch.view_as_real(eng.phi.A)).any()), "status": st.tolist(), "traj": traj}))
    print(json.dumps({"status_nonzero": int((st != 0).sum()), "values": sorted(set(st.tolist())), "stats": {k: int(v) for k, v in eng.stats.items()},
                      "events": len(events), "obs_nan": bool(torch.isnan(r).any())}))


if __name__ == "__main__":
    main()
