"""Micro-benchmark of the H_eff matvec (project_site) kernel chain at BASELINE sizes: the round-1 chain (RCONTRACT GEMM, k_wapply,
LCONTRACT GEMM) against the fused first step (yb_rcontract_mpo, 8- and 16-deep k tiles).  One JSON line: ms per kernel and the
fraction of the measured FP64 peak (profiles/r01_fp64_peak.json, 37.1 TFLOP/s) on the algorithmic flops."""
import json
import sys

import torch

sys.path.insert(0, ".")
from yaqs_b200 import kernels as K  # noqa: E402
from yaqs_b200.kernels import Buf, Ext  # noqa: E402

PEAK = 37.1


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
    dev = "cuda"
    for (nb, chi, D, dphys) in [(128, 128, 3, 4), (128, 113, 3, 4), (128, 128, 3, 2), (256, 64, 5, 4), (32, 256, 5, 4), (2048, 16, 3, 4)]:
        ext = torch.full((nb, 2), chi, dtype=torch.int32, device=dev)
        el, er = Ext(ext, 0, 2), Ext(ext, 1, 2)
        c2 = chi * chi
        x = torch.randn(nb * dphys * c2, dtype=torch.complex128, device=dev)
        renv = torch.randn(nb * D * c2, dtype=torch.complex128, device=dev)
        lenv = torch.randn(nb * D * c2, dtype=torch.complex128, device=dev)
        t1 = torch.empty(nb * dphys * D * c2, dtype=torch.complex128, device=dev)
        t2 = torch.empty(nb * dphys * D * c2, dtype=torch.complex128, device=dev)
        t3 = torch.empty(nb * dphys * D * c2, dtype=torch.complex128, device=dev)
        y = torch.empty(nb * dphys * c2, dtype=torch.complex128, device=dev)
        w = torch.randn(dphys * dphys * D * D, dtype=torch.complex128, device=dev)
        w[torch.rand(w.shape, device=dev) > 0.3] = 0
        bx, br, bl, b1, b2, b3, by = (Buf(x, 0, dphys * c2), Buf(renv, 0, D * c2), Buf(lenv, 0, D * c2), Buf(t1, 0, dphys * D * c2),
                                      Buf(t2, 0, dphys * D * c2), Buf(t3, 0, dphys * D * c2), Buf(y, 0, dphys * c2))
        f1 = lambda: kern.gemm(K.OP_RCONTRACT, nb, 1, None, el, er, None, dphys, D, D, bx, br, b1, 0, dphys * chi, D * chi)
        f2 = lambda: kern.wapply(nb, None, el, er, dphys, dphys, D, D, w, b1, b2, c2)
        f3 = lambda: kern.gemm(K.OP_LCONTRACT, nb, dphys, None, el, er, None, dphys, D, D, bl, b2, by, 0, chi, chi)
        ff = lambda: kern.rcontract_mpo(nb, None, el, er, dphys, D, D, bx, br, w, b3, chi, chi)
        fl = 8.0 * dphys * D * chi ** 3 * nb
        flw = 8.0 * dphys * dphys * D * D * chi * chi * nb
        m1, m2, m3 = timed(f1), timed(f2), timed(f3)
        mf8 = timed(ff)
        err8 = float((t3 - t2).abs().max().item())
        kern.set_option("gemm_warps16", 1)                      # 16 warps, 16 x 32 warp tiles
        mf16 = timed(ff)
        err16 = float((t3 - t2).abs().max().item())
        m1w, m3w = timed(f1), timed(f3)
        kern.set_option("gemm_warps16", 0)
        key = f"b{nb}_chi{chi}_D{D}_d{dphys}"
        out[key] = {"rcontract_ms": m1, "wapply_ms": m2, "lcontract_ms": m3, "fused_bk8_ms": mf8, "fused_bk16_ms": mf16,
                    "rcontract_warps16_ms": m1w, "lcontract_warps16_ms": m3w, "fused_warps16_ms": mf16,
                    "matvec_fused_warps16_frac": (2 * fl + flw) / (mf16 + m3w) / 1e9 / PEAK,
                    "rcontract_frac": fl / m1 / 1e9 / PEAK, "lcontract_frac": fl / m3 / 1e9 / PEAK,
                    "fused_bk8_frac": (fl + flw) / mf8 / 1e9 / PEAK, "fused_bk16_frac": (fl + flw) / mf16 / 1e9 / PEAK,
                    "matvec_round1_frac": (2 * fl + flw) / (m1 + m2 + m3) / 1e9 / PEAK, "matvec_fused_frac": (2 * fl + flw) / (min(mf8, mf16) + m3) / 1e9 / PEAK,
                    "wapply_gbs": 2 * 16.0 * nb * dphys * D * c2 / m2 / 1e6, "fused_vs_chain_max_abs_diff": max(err8, err16)}
        xa = x.view(nb, dphys * chi, chi); rb = renv.view(nb, chi, D * chi)
        ml = timed(lambda: torch.bmm(xa, rb))
        out[key]["cublas_rcontract_frac"] = fl / ml / 1e9 / PEAK
    print(json.dumps(out))


if __name__ == "__main__":
    main()
