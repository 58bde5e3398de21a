"""Micro-benchmark of the Hermitian eigen-solver: round-1 route (yb_heev + EVEC product) against yb_heev_rows (blocked, round 2)."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from yaqs_b200 import kernels as K  # noqa: E402
from yaqs_b200.kernels import Buf, Ext  # noqa: E402


def main():
    kern = K.CudaKernels("cuda:0")
    dev = "cuda:0"
    out = {}
    for (k, nb) in [(256, 128), (256, 64), (256, 148), (128, 128), (128, 64), (64, 128), (512, 74), (512, 16)]:
        sq = k * k
        rng = np.random.default_rng(k + nb)
        a = rng.standard_normal((k, k)) + 1j * rng.standard_normal((k, k))
        a *= np.logspace(0, -5, k)[None, :]
        g0 = torch.from_numpy(a @ a.conj().T).to(dev).reshape(-1).repeat(nb)
        ek = torch.full((nb,), k, dtype=torch.int32, device=dev)
        g = g0.clone()
        qh, zt, qa, eh = (torch.zeros(nb * sq, dtype=torch.complex128, device=dev) for _ in range(4))
        zr = torch.zeros(nb * sq, dtype=torch.float64, device=dev)
        d, e = (torch.zeros(nb * k, dtype=torch.float64, device=dev) for _ in range(2))
        tau = torch.zeros(nb * k, dtype=torch.complex128, device=dev)
        meta = torch.zeros(nb * 8, dtype=torch.int32, device=dev)
        status = torch.zeros(nb, dtype=torch.int32, device=dev)
        ext = Ext(ek, 0, 1)
        res = {}
        order = torch.zeros(nb * k, dtype=torch.int32, device=dev)
        ds = torch.zeros(nb * k, dtype=torch.float64, device=dev)
        ntop = torch.zeros(nb, dtype=torch.int32, device=dev)
        for name in ("round1", "rows", "rows_top"):
            for rep in range(3):
                g.copy_(g0)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                if name == "round1":
                    kern.heev(nb, None, ext, Buf(g, 0, sq), Buf(qh, 0, sq), Buf(zt, 0, sq), zr, sq, d, e, tau, k, meta, status, k)
                    kern.gemm(K.OP_EVEC, nb, 1, None, ext, None, None, 1, 1, 1, Buf(zt, 0, sq), Buf(qh, 0, sq), Buf(qa, 0, sq), 0, k, k)
                elif name == "rows_top":                       # capped split: only the k / 2 rows it can keep are back-transformed
                    kern.heev_rows_top(nb, None, ext, Buf(g, 0, sq), Buf(eh, 0, sq), zr, sq, d, e, tau, k, meta, status, k, k // 2, order, ds, ntop)
                else:
                    kern.heev_rows(nb, None, ext, Buf(g, 0, sq), Buf(eh, 0, sq), zr, sq, d, e, tau, k, meta, status, k)
                e1.record()
                torch.cuda.synchronize()
                res[name + "_ms"] = e0.elapsed_time(e1)
            if name == "rows_top":
                lam = ds[: k // 2].cpu().numpy().copy()
                E = eh[: (k // 2) * k].view(k // 2, k).cpu().numpy().conj().T
                gm = g0[:sq].view(k, k).cpu().numpy()
                res[name + "_resid"] = float(np.abs(gm @ E - E * lam[None, :]).max() / np.abs(lam).max())
                continue
            lam = d[:k].cpu().numpy().copy()
            E = (qa if name == "round1" else eh)[:sq].view(k, k).cpu().numpy().conj().T
            gm = g0[:sq].view(k, k).cpu().numpy()
            res[name + "_resid"] = float(np.abs(gm @ E - E * lam[None, :]).max() / np.abs(lam).max())
        res["status"] = int(status.max().item())
        out[f"k{k}_b{nb}"] = res
        print(f"k{k}_b{nb}", res, flush=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
