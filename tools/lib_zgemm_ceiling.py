"""Library (cuBLAS via torch) FP64 GEMM ceilings on the B200 box -- context numbers for the roofline.

Not part of the product path: it only records what a library batched ZGEMM/DGEMM reaches on the
shapes the TDVP contractions use, so DESIGN.md can state the FP64 denominator that was measured.
"""
import json
import sys
import time

import torch


def bench(fn, reps=10):
    fn(); fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    out = {"gpu": torch.cuda.get_device_name(0)}
    dev = "cuda"
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device=dev); b = torch.randn(n, n, dtype=torch.float64, device=dev)
    ms = bench(lambda: a @ b, 5)
    out["dgemm_8192_tflops"] = 2 * n**3 / ms / 1e9
    n = 4096
    a = torch.randn(n, n, dtype=torch.complex128, device=dev); b = torch.randn(n, n, dtype=torch.complex128, device=dev)
    ms = bench(lambda: a @ b, 5)
    out["zgemm_4096_tflops"] = 8 * n**3 / ms / 1e9
    # batched shapes of project_site at chi=128, D=3/4: (d~ chi) x chi x (D chi)
    for (bsz, m, k, nn) in [(512, 512, 128, 384), (512, 256, 128, 384), (512, 128, 384, 128), (512, 256, 128, 256),
                            (1024, 128, 64, 320), (2048, 32, 16, 48)]:
        a = torch.randn(bsz, m, k, dtype=torch.complex128, device=dev); b = torch.randn(bsz, k, nn, dtype=torch.complex128, device=dev)
        ms = bench(lambda: torch.bmm(a, b), 5)
        out[f"zbmm_b{bsz}_{m}x{k}x{nn}_tflops"] = 8.0 * bsz * m * k * nn / ms / 1e9
    # sustained 3 s loop for the power-capped figure
    n = 4096
    a = torch.randn(n, n, dtype=torch.complex128, device=dev); b = torch.randn(n, n, dtype=torch.complex128, device=dev)
    torch.cuda.synchronize(); t0 = time.time(); cnt = 0
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record()
    while time.time() - t0 < 3.0:
        for _ in range(5):
            a @ b
        cnt += 5
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    out["zgemm_4096_sustained_tflops"] = 8 * n**3 * cnt / e0.elapsed_time(e1) / 1e9
    print(json.dumps(out))


if __name__ == "__main__":
    sys.exit(main())
