"""Build libyaqs_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libyaqs_b200.so")
SOURCES = ["yb_api.cu", "yb_gemm.cu", "yb_ops.cu", "yb_svd.cu", "yb_heev.cu", "yb_stedc.cuh", "yb_philox.cuh", "yb_mpo_apply.cuh", "yb_wapply.cuh", "yb_common.cuh", "yb_truncate.cuh", "yb_cholqr.cuh", "yb_hetrd.cuh", "yb_gemm_rw.cuh"]
HEADER = os.path.join(os.path.dirname(HERE), "include", "yaqs_b200.h")


STAMP = LIB + ".sha256"


def source_digest() -> str:
    """Content hash of every source the library is built from (a checkout or a copy resets mtimes, contents do not lie)."""
    import hashlib
    h = hashlib.sha256()
    for path in sorted([os.path.join(CSRC, s) for s in os.listdir(CSRC) if s.endswith((".cu", ".cuh", ".h"))] + [HEADER]):
        h.update(os.path.basename(path).encode())
        with open(path, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def needs_build() -> bool:
    if not os.path.exists(LIB) or not os.path.exists(STAMP):
        return True
    with open(STAMP) as fh:
        return fh.read().strip() != source_digest()


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-shared", "-Xcompiler", "-fPIC",
           "-o", LIB, os.path.join(CSRC, "yb_api.cu")]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed:\n{res.stdout}\n{res.stderr}")
    if verbose:
        print(res.stderr)
    with open(STAMP, "w") as fh:
        fh.write(source_digest())
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
