"""Workload for ncu captures: the matvec kernels and the eigen-solver at C3 sizes, a few launches each."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from yaqs_b200 import kernels as K  # noqa: E402
from yaqs_b200.kernels import Buf, Ext  # noqa: E402


def main():
    kern = K.CudaKernels("cuda:0")
    dev = "cuda"
    nb, chi, D, dphys = 128, 128, 3, 4
    ext = torch.full((nb, 2), chi, dtype=torch.int32, device=dev)
    el, er = Ext(ext, 0, 2), Ext(ext, 1, 2)
    c2 = chi * chi
    x = torch.randn(nb * dphys * c2, dtype=torch.complex128, device=dev)
    renv = torch.randn(nb * D * c2, dtype=torch.complex128, device=dev)
    lenv = torch.randn(nb * D * c2, dtype=torch.complex128, device=dev)
    t1 = torch.empty(nb * dphys * D * c2, dtype=torch.complex128, device=dev)
    t2 = torch.empty(nb * dphys * D * c2, dtype=torch.complex128, device=dev)
    y = torch.empty(nb * dphys * c2, dtype=torch.complex128, device=dev)
    w = torch.randn(dphys * dphys * D * D, dtype=torch.complex128, device=dev)
    w[torch.rand(w.shape, device=dev) > 0.3] = 0
    bx, br, bl, b1, b2, by = (Buf(x, 0, dphys * c2), Buf(renv, 0, D * c2), Buf(lenv, 0, D * c2), Buf(t1, 0, dphys * D * c2),
                              Buf(t2, 0, dphys * D * c2), Buf(y, 0, dphys * c2))
    for _ in range(2):
        kern.gemm(K.OP_RCONTRACT, nb, 1, None, el, er, None, dphys, D, D, bx, br, b1, 0, dphys * chi, D * chi)
        kern.wapply(nb, None, el, er, dphys, dphys, D, D, w, b1, b2, c2)
        kern.gemm(K.OP_LCONTRACT, nb, dphys, None, el, er, None, dphys, D, D, bl, b2, by, 0, chi, chi)
        kern.rcontract_mpo(nb, None, el, er, dphys, D, D, bx, br, w, b2, chi, chi)
    k = 256
    sq = k * k
    rng = np.random.default_rng(1)
    a = rng.standard_normal((k, k)) + 1j * rng.standard_normal((k, k))
    a *= np.logspace(0, -5, k)[None, :]
    g0 = torch.from_numpy(a @ a.conj().T).to(dev).reshape(-1).repeat(nb)
    ek = torch.full((nb,), k, dtype=torch.int32, device=dev)
    eh = torch.zeros(nb * sq, dtype=torch.complex128, device=dev)
    zr = torch.zeros(nb * sq, dtype=torch.float64, device=dev)
    d, e = (torch.zeros(nb * k, dtype=torch.float64, device=dev) for _ in range(2))
    tau = torch.zeros(nb * k, dtype=torch.complex128, device=dev)
    meta = torch.zeros(nb * 8, dtype=torch.int32, device=dev)
    status = torch.zeros(nb, dtype=torch.int32, device=dev)
    for _ in range(2):
        g = g0.clone()
        kern.heev_rows(nb, None, Ext(ek, 0, 1), Buf(g, 0, sq), Buf(eh, 0, sq), zr, sq, d, e, tau, k, meta, status, k)
    # round 2, second session: the capped-split variant (top rows only) and the Cholesky-QR factorisation of the keep-all gauge shifts
    order = torch.zeros(nb * k, dtype=torch.int32, device=dev)
    ds = torch.zeros(nb * k, dtype=torch.float64, device=dev)
    ntop = torch.zeros(nb, dtype=torch.int32, device=dev)
    for _ in range(2):
        g = g0.clone()
        kern.heev_rows_top(nb, None, Ext(ek, 0, 1), Buf(g, 0, sq), Buf(eh, 0, sq), zr, sq, d, e, tau, k, meta, status, k, k // 2, order, ds, ntop)
    nb2, k2, len2 = 96, 128, 256
    a = rng.standard_normal((k2, len2)) + 1j * rng.standard_normal((k2, len2))
    a *= np.logspace(0, -3, k2)[:, None]
    w0 = torch.from_numpy(a).to(dev).reshape(-1).repeat(nb2)
    meta2 = torch.zeros(nb2 * 8, dtype=torch.int32, device=dev)
    meta2.view(nb2, 8)[:, 0] = k2
    meta2.view(nb2, 8)[:, 1] = len2
    g2 = torch.zeros(nb2 * k2 * k2, dtype=torch.complex128, device=dev)
    linv = torch.zeros(nb2 * k2 * k2, dtype=torch.complex128, device=dev)
    flag = torch.zeros(nb2, dtype=torch.int32, device=dev)
    nflag = torch.zeros(1, dtype=torch.int32, device=dev)
    ek2, el2 = Ext(meta2, 0, 8), Ext(meta2, 1, 8)
    for _ in range(2):
        kern.gemm(K.OP_GRAM, nb2, 1, None, ek2, el2, None, 1, 1, 1, Buf(w0, 0, k2 * len2), Buf(w0, 0, k2 * len2), Buf(g2, 0, k2 * k2), 0, k2, k2)
        kern.cholqr_factor(nb2, None, meta2, Buf(g2, 0, k2 * k2), Buf(linv, 0, k2 * k2), 1e-12, flag, nflag, k2)
    torch.cuda.synchronize()
    assert int(flag.max().item()) == 0


if __name__ == "__main__":
    main()
