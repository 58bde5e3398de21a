"""Micro-benchmark of the H_eff matvec (project_site) kernel chain at BASELINE config-3 sizes.

Prints one JSON line: per-kernel ms and the achieved fraction of the measured FP64 peak (profiles/r01_fp64_peak.json).
"""
import json
import sys

import torch

sys.path.insert(0, ".")
from yaqs_b200 import kernels as K  # noqa: E402
from yaqs_b200.kernels import Buf, Ext  # noqa: E402


def timed(fn, reps=5):
    fn(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    kern = K.CudaKernels("cuda:0")
    out = {}
    for (nb, chi, D, dphys) in [(256, 128, 3, 4), (256, 128, 3, 2), (256, 64, 5, 4), (64, 256, 5, 4), (2048, 16, 3, 4)]:
        dev = "cuda"
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
        f1 = lambda: kern.gemm(K.OP_RCONTRACT, nb, 1, None, el, er, None, dphys, D, D, Buf(x, 0, dphys * c2), Buf(renv, 0, D * c2),
                               Buf(t1, 0, dphys * D * c2), 0, dphys * chi, D * chi)
        f2 = lambda: kern.wapply(nb, None, el, er, dphys, dphys, D, D, w, Buf(t1, 0, dphys * D * c2), Buf(t2, 0, dphys * D * c2), c2)
        f3 = lambda: kern.gemm(K.OP_LCONTRACT, nb, dphys, None, el, er, None, dphys, D, D, Buf(lenv, 0, D * c2), Buf(t2, 0, dphys * D * c2),
                               Buf(y, 0, dphys * c2), 0, chi, chi)
        m1, m2, m3 = timed(f1), timed(f2), timed(f3)
        fl = 8.0 * dphys * D * chi ** 3 * nb
        key = f"b{nb}_chi{chi}_D{D}_d{dphys}"
        out[key] = {"rcontract_ms": m1, "wapply_ms": m2, "lcontract_ms": m3, "rcontract_tflops": fl / m1 / 1e9, "lcontract_tflops": fl / m3 / 1e9,
                    "matvec_tflops": (2 * fl + 8.0 * dphys * dphys * D * D * chi * chi * nb) / (m1 + m2 + m3) / 1e9,
                    "wapply_gbs": 2 * 16.0 * nb * dphys * D * c2 / m2 / 1e6}
        # library reference for the same two GEMMs
        xa = x.view(nb, dphys * chi, chi); rb = renv.view(nb, chi, D * chi)
        ml = timed(lambda: torch.bmm(xa, rb))
        out[key]["cublas_rcontract_tflops"] = fl / ml / 1e9
    print(json.dumps(out))


if __name__ == "__main__":
    main()
