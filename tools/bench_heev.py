"""Micro-benchmark of the Hermitian eigen-solver stages (run under ncu for per-kernel times, or alone for totals)."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from yaqs_b200 import kernels as K
from yaqs_b200.kernels import Buf, Ext


def main():
    kern = K.CudaKernels("cuda:0")
    dev = "cuda:0"
    out = {}
    for ring, (k, nb) in [(r, c) for r in (0, 6, 12, 16) for c in [(256, 8), (256, 128), (256, 296), (128, 296), (512, 74)]]:
        kern.set_option("heev_ring", ring)
        sq = k * k
        rng = np.random.default_rng(k + nb)
        a = rng.standard_normal((k, k)) + 1j * rng.standard_normal((k, k))
        a *= np.logspace(0, -5, k)[None, :]
        g0 = torch.from_numpy(a @ a.conj().T).to(dev).reshape(-1)
        ek = torch.full((nb,), k, dtype=torch.int32, device=dev)
        g = g0.repeat(nb)
        qh, zt, qa = (torch.zeros(nb * sq, dtype=torch.complex128, device=dev) for _ in range(3))
        zr = torch.zeros(nb * sq, dtype=torch.float64, device=dev)
        d, e = (torch.zeros(nb * k, dtype=torch.float64, device=dev) for _ in range(2))
        tau = torch.zeros(nb * k, dtype=torch.complex128, device=dev)
        meta = torch.zeros(nb * 8, dtype=torch.int32, device=dev)
        status = torch.zeros(nb, dtype=torch.int32, device=dev)
        ext = Ext(ek, 0, 1)
        for rep in range(2):
            g.copy_(g0.repeat(nb))
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            kern.heev(nb, None, ext, Buf(g, 0, sq), Buf(qh, 0, sq), Buf(zt, 0, sq), zr, sq, d, e, tau, k, meta, status, k)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
        e0.record()
        kern.gemm(K.OP_EVEC, nb, 1, None, ext, None, None, 1, 1, 1, Buf(zt, 0, sq), Buf(qh, 0, sq), Buf(qa, 0, sq), 0, k, k)
        e1.record()
        torch.cuda.synchronize()
        out[f"k{k}_b{nb}_ring{ring}"] = {"heev_ms": ms, "evec_gemm_ms": e0.elapsed_time(e1), "status": int(status.max().item())}
    kern.set_option("heev_ring", -1)
    import json
    print(json.dumps(out))


if __name__ == "__main__":
    main()
