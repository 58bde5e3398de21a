"""Diagnostic (GPU): replay the one-sided Jacobi schedule on the rank-deficient matrices that did not converge in the C3 run
(tests/golden/jacobi/rankdef_c3_*.npy) and print the rotation count of every sweep for several panel geometries."""
import glob
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from yaqs_b200.kernels import SVD_META, SVD_ROWS, Buf, CudaKernels, Ext  # noqa: E402


def run(kern, w0, sd, rows, nb, inner, sweeps=40, dump=None):
    dev = kern.device
    k, ln = w0.shape
    B = 1
    src = torch.zeros(sd * sd, dtype=torch.complex128, device=dev)
    src[: k * ln] = torch.from_numpy(w0.reshape(-1)).to(dev)          # mode SVD_ROWS with dphys = 1: (l) x r contiguous
    el = torch.tensor([k], dtype=torch.int32, device=dev)
    er = torch.tensor([ln], dtype=torch.int32, device=dev)
    w = torch.zeros(sd * sd, dtype=torch.complex128, device=dev)
    q = torch.zeros(sd * sd, dtype=torch.complex128, device=dev)
    meta = torch.zeros(B * SVD_META, dtype=torch.int32, device=dev)
    rot = torch.zeros(B, dtype=torch.int32, device=dev)
    counter = torch.zeros(1, dtype=torch.int32, device=dev)
    kern.svd_prepare(B, None, SVD_ROWS, Ext(el, 0, 1), Ext(er, 0, 1), 1, Buf(src, 0, sd * sd), Buf(w, 0, sd * sd), Buf(q, 0, sd * sd), meta, rot)
    hist = []
    for s in range(sweeps):
        for rnd in range(nb - 1):
            kern.svd_round(B, None, meta, rows, nb, rnd, rnd == 0, inner, Buf(w, 0, sd * sd), Buf(q, 0, sd * sd), rot, sd, sd)
        hist.append(int(rot.item()))
        if dump and s in dump:
            m = meta.cpu().numpy()
            np.save(dump[s], w[: int(m[0]) * int(m[1])].cpu().numpy().reshape(int(m[0]), int(m[1])))
        counter.zero_()
        kern.svd_check(B, None, meta, rot, counter)
        if int(counter.item()) == 0:
            break
    return hist, meta.cpu().numpy().tolist()


def main():
    kern = CudaKernels("cuda:0")
    out = {}
    for f in sorted(glob.glob("tests/golden/jacobi/rankdef_c3_*.npy")):
        w0 = np.load(f).astype(np.complex128)
        tag = f.split("_")[-1][:-4]
        for name, (sd, rows, nb, inner) in {"sd256_rows8_nb32": (256, 8, 32, 1), "sd256_rows16_nb16": (256, 16, 16, 1), "sd32_single_panel": (32, 16, 2, 30),
                                            "sd64_rows16_nb4": (64, 16, 4, 1), "sd256_rows4_nb64": (256, 4, 64, 1)}.items():
            dump = {25: f"gpurun_out/r02c2/w_{tag}_{name}_s25.npy", 26: f"gpurun_out/r02c2/w_{tag}_{name}_s26.npy"} if name == "sd256_rows8_nb32" else None
            hist, meta = run(kern, w0, sd, rows, nb, inner, dump=dump)
            out[f"{tag}:{name}"] = {"rotations_per_sweep": hist, "meta": meta}
            print(tag, name, hist, flush=True)
    print(json.dumps(out))


if __name__ == "__main__":
    import os
    os.makedirs("gpurun_out/r02c2", exist_ok=True)
    main()
