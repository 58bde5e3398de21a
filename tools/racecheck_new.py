"""Tiny runs of the kernels of the second session of round 2 for `compute-sanitizer` (racecheck: shared-memory hazards of
k_cholqr_factor, k_ormtr_rows in top mode, k_heev_order, k_jump_test; memcheck: the new global-memory index arithmetic)."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from yaqs_b200 import kernels as K  # noqa: E402
from yaqs_b200.kernels import Buf, Ext  # noqa: E402


def main():
    kern = K.CudaKernels("cuda:0")
    dev = "cuda:0"
    rng = np.random.default_rng(0)
    # ---- Cholesky-QR factorisation at several sizes (odd sizes exercise the partial quarter-warps)
    for k, ln, nb in [(5, 9, 2), (37, 70, 2), (128, 256, 1)]:
        a = rng.standard_normal((k, ln)) + 1j * rng.standard_normal((k, ln))
        gm = a @ a.conj().T
        g = torch.from_numpy(gm).to(dev).reshape(-1).repeat(nb)
        meta = torch.zeros(nb * 8, dtype=torch.int32, device=dev)
        meta.view(nb, 8)[:, 0] = k
        meta.view(nb, 8)[:, 1] = ln
        linv = torch.zeros(nb * k * k, dtype=torch.complex128, device=dev)
        flag = torch.zeros(nb, dtype=torch.int32, device=dev)
        nflag = torch.zeros(1, dtype=torch.int32, device=dev)
        kern.cholqr_factor(nb, None, meta, Buf(g, 0, k * k), Buf(linv, 0, k * k), 1e-12, flag, nflag, k)
        torch.cuda.synchronize()
        li = linv[: k * k].view(k, k).cpu().numpy()
        err = np.max(np.abs(li @ np.linalg.cholesky(gm) - np.eye(k)))
        print(f"cholqr k={k} flags={flag.cpu().tolist()} |Linv L - 1|={err:.2e}")
    # ---- eigen-solver in top mode
    for k, nb in [(70, 2), (256, 1)]:
        sq = k * k
        a = rng.standard_normal((k, k)) + 1j * rng.standard_normal((k, k))
        a *= np.logspace(0, -4, k)[None, :]
        gm = a @ a.conj().T
        g = torch.from_numpy(gm).to(dev).reshape(-1).repeat(nb)
        ek = torch.full((nb,), k, dtype=torch.int32, device=dev)
        eh = torch.zeros(nb * sq, dtype=torch.complex128, device=dev)
        zr = torch.zeros(nb * sq, dtype=torch.float64, device=dev)
        d, e, ds = (torch.zeros(nb * k, dtype=torch.float64, device=dev) for _ in range(3))
        order = torch.zeros(nb * k, dtype=torch.int32, device=dev)
        ntop = torch.zeros(nb, dtype=torch.int32, device=dev)
        tau = torch.zeros(nb * k, dtype=torch.complex128, device=dev)
        meta = torch.zeros(nb * 8, dtype=torch.int32, device=dev)
        status = torch.zeros(nb, dtype=torch.int32, device=dev)
        kern.heev_rows_top(nb, None, Ext(ek, 0, 1), Buf(g, 0, sq), Buf(eh, 0, sq), zr, sq, d, e, tau, k, meta, status, k, k // 2, order, ds, ntop)
        torch.cuda.synchronize()
        lam = ds[: k // 2].cpu().numpy()
        E = eh[: (k // 2) * k].view(k // 2, k).cpu().numpy().conj().T
        print(f"heev_rows_top k={k} status={int(status.max().item())} resid={np.abs(gm @ E - E * lam[None, :]).max() / lam[0]:.2e}")
    # ---- jump test / selection, observables, diagnostics
    B, L, nproc = 9, 4, 5
    out = {n: torch.zeros(B, dtype=torch.float64, device=dev) for n in ("dp", "u1", "u2")}
    jm, nj = torch.zeros(B, dtype=torch.int32, device=dev), torch.zeros(1, dtype=torch.int32, device=dev)
    pos = torch.arange(B, dtype=torch.int64, device=dev)
    traj = torch.arange(B, dtype=torch.int64, device=dev)
    kern.jump_test(B, torch.full((B,), 0.5, dtype=torch.float64, device=dev), None, 0, pos, (3, traj, 0x5452414A, 0), out["dp"], out["u1"], out["u2"], jm, nj)
    rdm1 = torch.ones(L * B * 4, dtype=torch.complex128, device=dev)
    w = torch.zeros(B * nproc, dtype=torch.float64, device=dev)
    tot = torch.zeros(B, dtype=torch.float64, device=dev)
    ch = torch.zeros(B, dtype=torch.int32, device=dev)
    sel = torch.zeros(B * 24, dtype=torch.complex128, device=dev)
    status = torch.zeros(B, dtype=torch.int32, device=dev)
    kern.jump_select(B, jm, out["u2"], rdm1, None, nproc, torch.zeros(nproc, dtype=torch.int32, device=dev), torch.zeros(nproc, dtype=torch.int32, device=dev),
                     torch.ones(nproc * 16, dtype=torch.complex128, device=dev), torch.ones(nproc, dtype=torch.float64, device=dev), 0.1,
                     torch.ones(nproc * 24, dtype=torch.complex128, device=dev), 24, w, tot, ch, sel, status)
    res = torch.zeros(B * 3, dtype=torch.float64, device=dev)
    kern.observables(B, 3, torch.arange(3, dtype=torch.int32, device=dev), torch.arange(3, dtype=torch.int32, device=dev), torch.zeros(3, dtype=torch.int32, device=dev),
                     torch.ones(3 * 16, dtype=torch.complex128, device=dev), rdm1, None, res, 3, status)
    bond = torch.full((B, L + 1), 3, dtype=torch.int32, device=dev)
    dg = torch.zeros(B * 3, dtype=torch.float64, device=dev)
    kern.diagnostics(B, L, 2, bond, dg, 3)
    torch.cuda.synchronize()
    print("jump njump", int(nj.item()), "choice", ch.cpu().tolist(), "diag", dg[:3].cpu().tolist())


if __name__ == "__main__":
    main()
