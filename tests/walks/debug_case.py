"""Debug helper: run one golden case on the CUDA engine and on the numpy twin side by side and report the first
high-level engine call after which the two states differ (gauge-invariant comparison through the dense state vector)."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from oracle.np_kernels import NumpyKernels  # noqa: E402
from oracle import tjm_oracle as O  # noqa: E402
from tests.helpers import engine_for_case  # noqa: E402
from yaqs_b200.kernels import CudaKernels  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "ising6_twosite_order2"
eg, tug, sug, _ = engine_for_case(name, CudaKernels("cuda:0"))
ec, tuc, suc, _ = engine_for_case(name, NumpyKernels())

TRACKED = ["shift_right", "shift_left", "split_theta", "normalize_site", "apply_dissipation", "tdvp", "_apply_jumps", "merge"]


def vec(eng, st, b):
    return O.mps_to_vector(O.OracleMPS(st.download(b)))


calls = {"n": 0}


def wrap(method):
    fg, fc = getattr(eg, method), getattr(ec, method)

    def run_both(*a, **k):
        raise RuntimeError

    return fg, fc


def compare(tag):
    for st_g, st_c, nm in ((eg.phi, ec.phi, "phi"), (eg.psi, ec.psi, "psi")):
        if st_g is None:
            continue
        bg, bc = st_g.bond.cpu().numpy(), st_c.bond.numpy()
        if not np.array_equal(bg, bc):
            print(f"[{tag}] {nm}: bond dims differ\n gpu {bg[0]}\n cpu {bc[0]}")
            return False
        for b in range(min(2, eg.B)):
            vg, vc = vec(eg, st_g, b), vec(ec, st_c, b)
            ov = abs(np.vdot(vg, vc))
            ng, nc = np.linalg.norm(vg), np.linalg.norm(vc)
            if abs(ng - nc) > 1e-9 or abs(ov - ng * nc) > 1e-9:
                print(f"[{tag}] {nm} traj {b}: |gpu|={ng:.12f} |cpu|={nc:.12f} overlap={ov:.12f}")
                return False
    return True


# monkeypatch the tracked methods so that each call is followed by a comparison
for m in ("shift_right", "shift_left", "split_theta", "normalize_site"):
    for eng in (eg, ec):
        orig = getattr(eng, m)

        def patched(*a, _orig=orig, _m=m, _eng=eng, **k):
            r = _orig(*a, **k)
            _eng._trace.append((_m, a[1] if len(a) > 1 else None))
            return r

        setattr(eng, m, patched)
    eg._trace, ec._trace = [], []

p = eg.p
steps = [("dissipation dt/2", lambda e: e.apply_dissipation(e.phi, p.dt / 2))]
for tag, fn in steps:
    # run op-by-op: execute on cpu fully to get the trace, then replay on gpu comparing after each traced op
    fn(ec)
    trace_c = list(ec._trace)
    # re-create the cpu engine to replay in lock-step
    ec2, _, _, _ = engine_for_case(name, NumpyKernels())
    ok = True
    lay, n = eg.lay, eg.L
    shift = eg._trunc_args("discarded_weight", 1e-12, None, 1)
    user = eg._trunc_args(p.trunc_mode, p.svd_threshold, p.max_bond_dim, 1)
    facs = eg._dissipators(p.dt / 2)
    engines = (eg, ec2)

    def both(desc, f):
        for e in engines:
            f(e)
        global ec
        ec = ec2
        good = compare(desc)
        print(desc, "OK" if good else "MISMATCH")
        return good

    for i in range(n - 1):
        if not both(f"shift_right {i}", lambda e, i=i: e.shift_right(e.phi, i, shift)):
            sys.exit(0)
    for i in reversed(range(n)):
        op1, scalar2, pair = facs[i]
        if op1 is not None:
            def f(e, i=i, op1=op1):
                op_t = torch.from_numpy(np.ascontiguousarray(op1).reshape(-1)).to(e.dev)
                e.K.apply_local_op(e.B, None, None, 0, e.phi.ext(i), e.phi.ext(i + 1), 2, op_t, 0, e.phi.site(i), lay.cap[i] * lay.cap[i + 1])
            if not both(f"op1 {i}", f):
                sys.exit(0)
        if scalar2 is not None:
            def f(e, i=i, s=scalar2):
                op_t = torch.from_numpy((s * np.eye(2, dtype=np.complex128)).reshape(-1)).to(e.dev)
                e.K.apply_local_op(e.B, None, None, 0, e.phi.ext(i), e.phi.ext(i + 1), 2, op_t, 0, e.phi.site(i), lay.cap[i] * lay.cap[i + 1])
            if not both(f"scalar2 {i}", f):
                sys.exit(0)
        if pair is not None:
            def f(e, i=i, pair=pair):
                th = e.merge(e.phi, i - 1)
                op_t = torch.from_numpy(np.ascontiguousarray(pair).reshape(-1)).to(e.dev)
                e.K.apply_local_op(e.B, None, None, 0, e.phi.ext(i - 1), e.phi.ext(i + 1), 4, op_t, 0, th, lay.cap[i - 1] * lay.cap[i + 1])
                nth = 4 * int(e.phi.bond[0, i - 1]) * int(e.phi.bond[0, i + 1])
                print("   theta", e.K.name, np.round(e.theta[:nth].cpu().numpy(), 6))
                e.split_theta(e.phi, i - 1, "right", user)
                print("   keep", e.keep.cpu().numpy()[:3], "sval", np.round(e.svd_s[:4].cpu().numpy(), 8))
            if not both(f"pair {i}", f):
                sys.exit(0)
        if i != 0:
            if not both(f"shift_left {i}", lambda e, i=i: e.shift_left(e.phi, i, shift)):
                sys.exit(0)
print("done")
