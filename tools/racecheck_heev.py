"""Tiny eigen-solver run for `compute-sanitizer --tool racecheck` (shared-memory hazards of k_hetrd / k_ungtr / k_stedc)."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from yaqs_b200 import kernels as K
from yaqs_b200.kernels import Buf, Ext


def main():
    kern = K.CudaKernels("cuda:0")
    dev = "cuda:0"
    for k, nb in [(70, 2), (128, 1), (256, 1)]:
        sq = k * k
        rng = np.random.default_rng(k)
        a = rng.standard_normal((k, k)) + 1j * rng.standard_normal((k, k))
        a *= np.logspace(0, -6, k)[None, :]
        gm = a @ a.conj().T
        g = torch.from_numpy(gm).to(dev).reshape(-1).repeat(nb)
        ek = torch.full((nb,), k, dtype=torch.int32, device=dev)
        qh, zt = (torch.zeros(nb * sq, dtype=torch.complex128, device=dev) for _ in range(2))
        zr = torch.zeros(nb * sq, dtype=torch.float64, device=dev)
        d, e = (torch.zeros(nb * k, dtype=torch.float64, device=dev) for _ in range(2))
        tau = torch.zeros(nb * k, dtype=torch.complex128, device=dev)
        meta = torch.zeros(nb * 8, dtype=torch.int32, device=dev)
        status = torch.zeros(nb, dtype=torch.int32, device=dev)
        kern.heev(nb, None, Ext(ek, 0, 1), Buf(g, 0, sq), Buf(qh, 0, sq), Buf(zt, 0, sq), zr, sq, d, e, tau, k, meta, status, k)
        torch.cuda.synchronize()
        lam = np.sort(d.cpu().numpy()[:k])
        ref = np.linalg.eigvalsh(gm)
        print(f"k={k} nb={nb} status={int(status.max().item())} max|dlam|={np.max(np.abs(lam - ref)):.3e}")


if __name__ == "__main__":
    main()
