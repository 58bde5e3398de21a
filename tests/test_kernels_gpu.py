"""-m gpu: every CUDA kernel behind the C ABI vs. its numpy twin (oracle/np_kernels.py) on seeded ragged inputs.

complex128 tolerance: 1e-12 relative to the operand norms for contractions (fp64 summation-order differences only);
SVD outputs are compared through gauge-invariant quantities (singular values, keep counts, reconstruction, isometry).
"""
import numpy as np
import pytest
import torch

from oracle.np_kernels import NumpyKernels
from yaqs_b200 import kernels as K
from yaqs_b200.kernels import Buf, Ext

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def backends():
    return K.CudaKernels("cuda:0"), NumpyKernels()


def _rand_c(rng, n):
    return torch.from_numpy(rng.standard_normal(n) + 1j * rng.standard_normal(n))


def _pair(t):
    return t.clone(), t.clone().cuda()


def _cmp(cpu, gpu, tol=1e-12):
    a, b = cpu.numpy(), gpu.cpu().numpy()
    scale = max(1.0, float(np.abs(a).max()))
    assert np.max(np.abs(a - b)) <= tol * scale, float(np.max(np.abs(a - b)))


GEMM_CASES = [
    # (op, ngroups, dphys, Dl, Dr, conj, sizes of A/B/C as functions of (x0, x1, x2))
    (K.OP_RCONTRACT, 1, 4, 3, 3, 0), (K.OP_RCONTRACT, 1, 2, 5, 4, 0), (K.OP_RCONTRACT, 1, 1, 3, 3, 0),
    (K.OP_LCONTRACT, 4, 4, 3, 3, 0), (K.OP_LCONTRACT, 2, 2, 4, 5, 0), (K.OP_LCONTRACT, 1, 1, 3, 3, 0),
    (K.OP_ENVR3, 1, 2, 3, 4, 0), (K.OP_ENVL1, 2, 2, 3, 4, 0), (K.OP_ENVL3, 1, 2, 3, 4, 0), (K.OP_ENVL3, 1, 2, 1, 1, 1),
    (K.OP_MERGE, 4, 2, 1, 1, 0), (K.OP_ABSORB_R, 2, 2, 1, 1, 0), (K.OP_ABSORB_L, 2, 2, 1, 1, 0),
    (K.OP_GRAM, 1, 1, 1, 1, 0), (K.OP_EVEC, 1, 1, 1, 1, 0), (K.OP_ROTATE, 1, 1, 1, 1, 0),
]


def _gemm_sizes(op, ngroups, dphys, Dl, Dr, cap):
    c2 = cap * cap
    if op == K.OP_RCONTRACT:
        return dphys * c2, Dr * c2, dphys * Dr * c2
    if op == K.OP_LCONTRACT:
        return Dl * c2, dphys * Dl * c2, dphys * c2
    if op == K.OP_ENVR3:
        return dphys * Dl * c2, dphys * c2, Dl * c2
    if op == K.OP_ENVL1:
        return Dl * c2, dphys * c2, dphys * Dl * c2
    if op == K.OP_ENVL3:
        return dphys * c2, dphys * Dr * c2, Dr * c2
    if op == K.OP_MERGE:
        return dphys * c2, dphys * c2, dphys * dphys * c2
    return ngroups * c2, ngroups * c2, ngroups * c2


@pytest.mark.parametrize("case", GEMM_CASES)
@pytest.mark.parametrize("cap", [5, 37, 96])
def test_gemm_ops_match_twin(backends, case, cap):
    cuda, twin = backends
    op, ngroups, dphys, Dl, Dr, conj = case
    rng = np.random.default_rng(1000 * op + cap)
    nb = 5
    ext = torch.from_numpy(rng.integers(1, cap + 1, size=(nb, 3)).astype(np.int32))
    ext[0] = cap                                  # one trajectory at full capacity
    mask = torch.tensor([1, 1, 0, 1, 1], dtype=torch.int32)
    sa, sb, sc = _gemm_sizes(op, ngroups, dphys, Dl, Dr, cap)
    a, b = _rand_c(rng, nb * sa + 7), _rand_c(rng, nb * sb + 3)
    c = _rand_c(rng, nb * sc)
    ext_c, ext_g = _pair(ext)
    mask_c, mask_g = _pair(mask)
    a_c, a_g = _pair(a)
    b_c, b_g = _pair(b)
    c_c, c_g = _pair(c)
    maxm = 4 * 5 * cap
    for kern, e, m, aa, bb, cc in ((twin, ext_c, mask_c, a_c, b_c, c_c), (cuda, ext_g, mask_g, a_g, b_g, c_g)):
        kern.gemm(op, nb, ngroups, m, Ext(e, 0, 3), Ext(e, 1, 3), Ext(e, 2, 3), dphys, Dl, Dr, Buf(aa, 7, sa), Buf(bb, 3, sb),
                  Buf(cc, 0, sc), conj, maxm, maxm)
    torch.cuda.synchronize()
    _cmp(c_c, c_g, 1e-12 * cap)


@pytest.mark.parametrize("dims", [(4, 4, 3, 3), (2, 2, 5, 4), (2, 2, 1, 1), (4, 4, 5, 5)])
def test_wapply_matches_twin(backends, dims):
    cuda, twin = backends
    nx, nu, ny, nv = dims
    rng = np.random.default_rng(7)
    nb, cap = 4, 19
    ext = torch.from_numpy(rng.integers(1, cap + 1, size=(nb, 2)).astype(np.int32))
    w = _rand_c(rng, nx * nu * ny * nv)
    w[::3] = 0                                        # sparse MPO tensors: zero entries are skipped
    sin, sout = nu * nv * cap * cap, nx * ny * cap * cap
    src, dst = _rand_c(rng, nb * sin), _rand_c(rng, nb * sout)
    ext_c, ext_g = _pair(ext)
    w_c, w_g = _pair(w)
    s_c, s_g = _pair(src)
    d_c, d_g = _pair(dst)
    for kern, e, ww, ss, dd in ((twin, ext_c, w_c, s_c, d_c), (cuda, ext_g, w_g, s_g, d_g)):
        kern.wapply(nb, None, Ext(e, 0, 2), Ext(e, 1, 2), nx, nu, ny, nv, ww, Buf(ss, 0, sin), Buf(dd, 0, sout), cap * cap)
    torch.cuda.synchronize()
    _cmp(d_c, d_g)


@pytest.mark.parametrize("warps16", [0, 1])
@pytest.mark.parametrize("dims", [(4, 3, 3, 40), (4, 4, 3, 128), (2, 3, 3, 37), (2, 5, 5, 64), (4, 5, 4, 33), (1, 3, 3, 17), (4, 1, 1, 21), (2, 4, 1, 8),
                                  (4, 5, 5, 130)])
def test_rcontract_mpo_matches_twin(backends, dims, warps16):
    """project_site's first two contractions in one kernel (MPO tensor applied in the GEMM epilogue) against the numpy twin --
    ragged extents, both warp layouts, MPO tensors with the reference's sparsity (~30 % non-zero)."""
    cuda, twin = backends
    dphys, dl, dr, cap = dims
    rng = np.random.default_rng(11 * cap + dphys + dl)
    nb = 5
    ext = rng.integers(1, cap + 1, size=(nb, 2)).astype(np.int32)
    ext[0] = cap
    ext[1] = (cap, 1)
    ext[2] = (1, cap)
    mask = np.array([1, 1, 1, 0, 1], dtype=np.int32)
    sx, sr, so = dphys * cap * cap, cap * dr * cap, dphys * cap * dl * cap
    rc = lambda n: rng.standard_normal(n) + 1j * rng.standard_normal(n)      # noqa: E731
    x, r, w = rc(nb * sx), rc(nb * sr), rc(dphys * dphys * dl * dr)
    w[rng.random(w.shape) > 0.35] = 0
    outs = []
    cuda.set_option("gemm_warps16", warps16)
    try:
        for kern, dev in ((twin, "cpu"), (cuda, "cuda")):
            et = torch.from_numpy(ext).to(dev)
            out = torch.full((nb * so,), 7.0 + 0j, dtype=torch.complex128, device=dev)
            kern.rcontract_mpo(nb, torch.from_numpy(mask).to(dev), Ext(et, 0, 2), Ext(et, 1, 2), dphys, dl, dr, Buf(torch.from_numpy(x).to(dev), 0, sx),
                               Buf(torch.from_numpy(r).to(dev), 0, sr), torch.from_numpy(w).to(dev), Buf(out, 0, so), cap, cap)
            outs.append(out.cpu().numpy().reshape(nb, so))
    finally:
        cuda.set_option("gemm_warps16", 0)
    for t in range(nb):
        n = dphys * ext[t, 0] * dl * ext[t, 1] if mask[t] else 0
        scale = max(1.0, np.abs(outs[0][t, :n]).max()) if n else 1.0
        assert np.max(np.abs(outs[0][t, :n] - outs[1][t, :n]), initial=0.0) < 1e-12 * scale * cap
        assert np.all(outs[1][t, n:] == 7.0)                     # nothing written beyond the packed tensor / for masked trajectories


def test_small_utilities_match_twin(backends):
    cuda, twin = backends
    rng = np.random.default_rng(11)
    nb, cap, d = 6, 13, 2
    ext = torch.from_numpy(rng.integers(1, cap + 1, size=(nb, 2)).astype(np.int32))
    mask = torch.tensor([1, 0, 1, 1, 1, 1], dtype=torch.int32)
    sel = torch.tensor([3, 3, 2, 3, 1, 3], dtype=torch.int32)
    slot = d * cap * cap
    a = _rand_c(rng, nb * slot)
    x = _rand_c(rng, nb * slot)
    ops = _rand_c(rng, nb * d * d)
    scale = torch.from_numpy(rng.standard_normal(nb))
    res = {}
    for name, kern, dev in (("cpu", twin, "cpu"), ("gpu", cuda, "cuda")):
        e, m, s = ext.to(dev), mask.to(dev), sel.to(dev)
        aa, xx, oo, sc = a.clone().to(dev), x.clone().to(dev), ops.to(dev), scale.to(dev)
        el, er = Ext(e, 0, 2), Ext(e, 1, 2)
        out = {}
        n2 = torch.zeros(nb, dtype=torch.float64, device=dev)
        kern.norm2(nb, m, el, er, d, Buf(aa, 0, slot), n2)
        out["norm2"] = n2
        g = torch.zeros(nb * d * d, dtype=torch.complex128, device=dev)
        kern.local_rdm(nb, m, el, er, d, Buf(aa, 0, slot), Buf(xx, 0, slot), g)
        out["rdm"] = g
        kern.apply_local_op(nb, m, s, 3, el, er, d, oo, d * d, Buf(aa, 0, slot), cap * cap)
        kern.apply_local_op(nb, None, None, 0, el, er, d, oo[: d * d].clone(), 0, Buf(xx, 0, slot), cap * cap)
        kern.scale(nb, m, el, er, d, sc, Buf(xx, 0, slot), slot)
        env = torch.zeros(nb * cap * 3 * cap, dtype=torch.complex128, device=dev)
        kern.env_identity(nb, m, el, 3, Buf(env, 0, cap * 3 * cap), cap)
        out["env"] = env
        un = torch.zeros(nb * slot, dtype=torch.complex128, device=dev)
        kern.unpack_vh(nb, m, el, er, d, Buf(aa, 0, slot), Buf(un, 0, slot), slot)
        cp = torch.zeros(nb * slot, dtype=torch.complex128, device=dev)
        kern.copy_packed(nb, m, el, er, d, Buf(xx, 0, slot), Buf(cp, 0, slot), slot)
        out.update(a=aa, x=xx, unpack=un, copy=cp)
        res[name] = out
    torch.cuda.synchronize()
    for key in res["cpu"]:
        _cmp(res["cpu"][key], res["gpu"][key])


def _run_lanczos(kern, dev, nb, n_list, hmats, x0, dt, tol, kmax=25):
    """Drive init/step/combine with a dense Hermitian matvec done on the host side of the test."""
    nmax = max(n_list)
    ext = torch.tensor([[1, n] for n in n_list], dtype=torch.int32, device=dev)
    el, er = Ext(ext, 0, 2), Ext(ext, 1, 2)
    svec = nmax
    basis = torch.zeros(nb * (kmax + 1) * svec, dtype=torch.complex128, device=dev)
    bb = Buf(basis, 0, (kmax + 1) * svec)
    x = torch.zeros(nb * nmax, dtype=torch.complex128, device=dev)
    for t, n in enumerate(n_list):
        x[t * nmax: t * nmax + n] = torch.from_numpy(x0[t]).to(dev)
    nrm = torch.zeros(nb, dtype=torch.float64, device=dev)
    active = torch.zeros(nb, dtype=torch.int32, device=dev)
    kdone = torch.zeros(nb, dtype=torch.int32, device=dev)
    alpha = torch.zeros(nb * 32, dtype=torch.float64, device=dev)
    beta = torch.zeros(nb * 32, dtype=torch.float64, device=dev)
    coef = torch.zeros(nb * 32, dtype=torch.complex128, device=dev)
    nact = torch.zeros(1, dtype=torch.int32, device=dev)
    kern.lanczos_init(nb, None, el, er, 1, Buf(x, 0, nmax), bb, svec, nrm, active, kdone)
    for j in range(kmax):
        bview = basis.view(nb, kmax + 1, svec)
        for t, n in enumerate(n_list):                      # matvec: w = H v_j into slot j+1
            if int(active[t]):
                h = torch.from_numpy(hmats[t]).to(dev)
                bview[t, j + 1, :n] = h @ bview[t, j, :n]
        nact.zero_()
        kern.lanczos_step(nb, j, kmax, dt, tol, el, er, 1, bb, svec, alpha, beta, nrm, active, kdone, coef, nact)
        if kern.sync_int(nact) == 0:
            break
    out = x.clone()
    kern.lanczos_combine(nb, None, el, er, 1, bb, svec, coef, kmax, kdone, Buf(out, 0, nmax))
    return out.cpu().numpy().reshape(nb, nmax), kdone.cpu().numpy()


def test_lanczos_matches_twin_and_expm(backends):
    import scipy.linalg
    cuda, twin = backends
    rng = np.random.default_rng(5)
    n_list = [40, 3, 64, 1, 17, 64]
    hm, xs = [], []
    for n in n_list:
        m = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
        hm.append((m + m.conj().T) / (2 * np.sqrt(n)))
        xs.append(rng.standard_normal(n) + 1j * rng.standard_normal(n))
    xs[3] = np.zeros(1, dtype=np.complex128)                 # zero vector is returned untouched
    dt, tol = 0.1, 1e-10
    out_g, k_g = _run_lanczos(cuda, "cuda", len(n_list), n_list, hm, xs, dt, tol)
    out_c, k_c = _run_lanczos(twin, "cpu", len(n_list), n_list, hm, xs, dt, tol)
    assert np.array_equal(k_g, k_c)
    for t, n in enumerate(n_list):
        assert np.max(np.abs(out_g[t, :n] - out_c[t, :n])) < 1e-12
        exact = scipy.linalg.expm(-1j * dt * hm[t]) @ xs[t]
        assert np.max(np.abs(out_g[t, :n] - exact)) < 1e-8


def _run_svd(kern, dev, mode, dphys, ext_np, src_np, slot, trunc, thr, cap_bond, min_keep, dist, block_rows=8):
    nb = ext_np.shape[0]
    ext = torch.from_numpy(ext_np).to(dev)
    el, er = Ext(ext, 0, 2), Ext(ext, 1, 2)
    cap = int(ext_np.max())
    maxdim = dphys * cap
    src = torch.from_numpy(src_np).to(dev)
    w = torch.zeros(nb * maxdim * maxdim, dtype=torch.complex128, device=dev)
    q = torch.zeros(nb * maxdim * maxdim, dtype=torch.complex128, device=dev)
    meta = torch.zeros(nb * 8, dtype=torch.int32, device=dev)
    rot = torch.zeros(nb, dtype=torch.int32, device=dev)
    pend = torch.zeros(1, dtype=torch.int32, device=dev)
    wb, qb = Buf(w, 0, maxdim * maxdim), Buf(q, 0, maxdim * maxdim)
    kern.svd_prepare(nb, None, mode, el, er, dphys, Buf(src, 0, slot), wb, qb, meta, rot)
    nblocks = max(2, 2 * ((maxdim + 2 * block_rows - 1) // (2 * block_rows)))
    for sweep in range(40):
        for rnd in range(nblocks - 1):
            kern.svd_round(nb, None, meta, block_rows, nblocks, rnd, rnd == 0, 1, wb, qb, rot, maxdim, maxdim)
        pend.zero_()
        kern.svd_check(nb, None, meta, rot, pend)
        if kern.sync_int(pend) == 0:
            break
    sval = torch.zeros(nb * maxdim, dtype=torch.float64, device=dev)
    keep = torch.zeros(nb, dtype=torch.int32, device=dev)
    status = torch.zeros(nb, dtype=torch.int32, device=dev)
    u = torch.zeros(nb * maxdim * maxdim, dtype=torch.complex128, device=dev)
    v = torch.zeros(nb * maxdim * maxdim, dtype=torch.complex128, device=dev)
    kern.svd_finalize(nb, None, meta, trunc, thr, cap_bond, min_keep, dist, wb, qb, sval, maxdim, Ext(keep, 0, 1),
                      Buf(u, 0, maxdim * maxdim), Buf(v, 0, maxdim * maxdim), status, maxdim)
    return (sval.cpu().numpy().reshape(nb, maxdim), keep.cpu().numpy(), u.cpu().numpy().reshape(nb, -1), v.cpu().numpy().reshape(nb, -1),
            meta.cpu().numpy().reshape(nb, 8), status.cpu().numpy(), sweep + 1)


@pytest.mark.parametrize("mode", [K.SVD_THETA, K.SVD_ROWS, K.SVD_COLS])
@pytest.mark.parametrize("dist", [0, 1, 2])
def test_svd_matches_lapack(backends, mode, dist):
    cuda, twin = backends
    rng = np.random.default_rng(100 + 10 * mode + dist)
    d, cap, nb = 2, 24, 6
    ext = rng.integers(1, cap + 1, size=(nb, 2)).astype(np.int32)
    ext[0] = cap
    ext[1] = (cap, 3)
    ext[2] = (2, cap)
    nphys = d * d if mode == K.SVD_THETA else d
    slot = nphys * cap * cap
    src = (rng.standard_normal(nb * slot) + 1j * rng.standard_normal(nb * slot))
    # graded spectrum for trajectory 3 so that the threshold actually truncates: low-rank + tiny noise
    l3, r3 = ext[3]
    args = (mode, d, ext, src, slot, 0, 1e-3, 12, 2, dist)
    sg, kg, ug, vg, mg, stg, nsweeps = _run_svd(cuda, "cuda", *args)
    sc, kc, uc, vc, mc, _, _ = _run_svd(twin, "cpu", *args)
    assert nsweeps < 30 and not stg.any()
    assert np.array_equal(kg, kc)
    for t in range(nb):
        k, rows, cols = mc[t, 0], mc[t, 3], mc[t, 4]
        assert np.max(np.abs(sg[t, :k] - sc[t, :k])) < 1e-12 * max(1.0, sc[t, 0])
        kp = kg[t]
        U_g, V_g = ug[t, : rows * kp].reshape(rows, kp), vg[t, : kp * cols].reshape(kp, cols)
        U_c, V_c = uc[t, : rows * kp].reshape(rows, kp), vc[t, : kp * cols].reshape(kp, cols)
        assert np.max(np.abs(U_g @ V_g - U_c @ V_c)) < 1e-11 * max(1.0, sc[t, 0])
        iso = (V_g @ V_g.conj().T) if dist == 0 else (U_g.conj().T @ U_g) if dist == 1 else None
        if iso is not None:
            assert np.max(np.abs(iso - np.eye(kp))) < 1e-12


def test_svd_rank_deficient_min_keep(backends):
    """Product-state theta: rank 1, min_keep = 2 forces a zero singular value into the kept set; the non-S factor must
    still be an isometry (LAPACK returns an arbitrary orthonormal completion there)."""
    cuda, _ = backends
    d, nb = 2, 3
    ext = np.array([[1, 1], [2, 3], [4, 4]], dtype=np.int32)
    cap = 4
    slot = d * d * cap * cap
    src = np.zeros(nb * slot, dtype=np.complex128)
    rng = np.random.default_rng(3)
    for t, (l, r) in enumerate(ext):
        a = rng.standard_normal(d * l) + 1j * rng.standard_normal(d * l)
        b = rng.standard_normal(d * r) + 1j * rng.standard_normal(d * r)
        theta = np.outer(a, b).reshape(d, l, d, r).transpose(0, 2, 1, 3)          # ((s s'), l, r)
        src[t * slot: t * slot + theta.size] = theta.reshape(-1)
    for dist in (0, 1):
        sg, kg, ug, vg, mg, stg, _ = _run_svd(cuda, "cuda", K.SVD_THETA, d, ext, src, slot, 0, 1e-9, 8, 2, dist)
        assert list(kg) == [2, 2, 2]
        for t, (l, r) in enumerate(ext):
            rows, cols, kp = d * l, d * r, 2
            U, V = ug[t, : rows * kp].reshape(rows, kp), vg[t, : kp * cols].reshape(kp, cols)
            theta = src[t * slot: t * slot + d * d * l * r].reshape(d, d, l, r).transpose(0, 2, 1, 3).reshape(rows, cols)
            assert np.max(np.abs(U @ V - theta)) < 1e-13 * np.abs(theta).max() + 1e-14
            iso = V @ V.conj().T if dist == 0 else U.conj().T @ U
            assert np.max(np.abs(iso - np.eye(kp))) < 1e-12
            assert sg[t, 1] < 1e-14 * sg[t, 0]


def _run_jacobi_on_matrix(kern, w0, pitch, block_rows, nblocks, inner, max_sweeps=40):
    """One matrix (k x len, rows to orthogonalise) through prepare / sweeps / finalize(keep all) with a given panel geometry.
    Returns (singular values, U, Vh, sweeps, rotations per sweep, status)."""
    dev = kern.device
    k, ln = w0.shape
    sq = pitch * pitch
    src = torch.zeros(sq, dtype=torch.complex128, device=dev)
    src[: k * ln] = torch.from_numpy(np.ascontiguousarray(w0).reshape(-1)).to(dev)
    el = torch.tensor([k], dtype=torch.int32, device=dev)
    er = torch.tensor([ln], dtype=torch.int32, device=dev)
    w, q, u, v = (torch.zeros(sq, dtype=torch.complex128, device=dev) for _ in range(4))
    meta = torch.zeros(8, dtype=torch.int32, device=dev)
    rot = torch.zeros(1, dtype=torch.int32, device=dev)
    pend = torch.zeros(1, dtype=torch.int32, device=dev)
    wb, qb = Buf(w, 0, sq), Buf(q, 0, sq)
    kern.svd_prepare(1, None, K.SVD_ROWS, Ext(el, 0, 1), Ext(er, 0, 1), 1, Buf(src, 0, sq), wb, qb, meta, rot)
    hist = []
    for sweep in range(max_sweeps):
        for rnd in range(nblocks - 1):
            kern.svd_round(1, None, meta, block_rows, nblocks, rnd, rnd == 0, inner, wb, qb, rot, pitch, pitch)
        hist.append(int(rot.item()))
        pend.zero_()
        kern.svd_check(1, None, meta, rot, pend)
        if kern.sync_int(pend) == 0:
            break
    sval = torch.zeros(pitch, dtype=torch.float64, device=dev)
    keep = torch.zeros(1, dtype=torch.int32, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    kern.svd_finalize(1, None, meta, 0, 0.0, 0, 1, 1, wb, qb, sval, pitch, Ext(keep, 0, 1), Buf(u, 0, sq), Buf(v, 0, sq), status, pitch)
    kp = int(keep.item())
    return (sval.cpu().numpy()[:k], u.cpu().numpy()[: k * kp].reshape(k, kp), v.cpu().numpy()[: kp * ln].reshape(kp, ln), sweep + 1, hist,
            int(status.item()), int(meta.cpu().numpy()[7]))


@pytest.mark.parametrize("which", [1, 2, 3])
@pytest.mark.parametrize("geometry", ["block8_of_256", "block4_of_256", "single_panel"])
def test_jacobi_converges_on_rank_deficient_post_jump_matrices(backends, golden_dir, which, geometry):
    """tests/golden/jacobi/rankdef_c3_*.npy: the three chain-end matrices (24 x 28, 14 x 16, 12 x 32; numerical rank 10 / 6 / 8, the
    rest rounding residue at 1e-19) on which the block schedule of round 1 still rotated after 60 sweeps -- 3 of 256 trajectories of
    the driver's C3 run ended with status bit 1.  With the per-matrix noise floor every geometry converges in a few sweeps; the
    factors reproduce the matrix, U is an isometry and the singular values above the floor agree with LAPACK."""
    cuda, _ = backends
    w0 = np.load(f"{golden_dir}/jacobi/rankdef_c3_{which}.npy").astype(np.complex128)
    k, ln = w0.shape
    pitch, rows, nb, inner = {"block8_of_256": (256, 8, 32, 1), "block4_of_256": (256, 4, 64, 1), "single_panel": (32, 16, 2, 30)}[geometry]
    sv, u, vh, sweeps, hist, status, scale_word = _run_jacobi_on_matrix(cuda, w0, pitch, rows, nb, inner)
    ref = np.linalg.svd(w0, compute_uv=False)
    assert status == 0 and sweeps <= 20, (sweeps, hist)
    assert scale_word == 0                                       # finalize hands the scale word back cleared
    floor = 20 * 2.3e-16 * ref[0]
    big = ref > floor
    assert np.max(np.abs(sv[: big.sum()] - ref[big])) < 1e-14 * ref[0]
    assert np.all(sv[big.sum():] <= floor)
    assert np.max(np.abs(u @ vh - w0)) < 1e-14 * ref[0]          # dist = right: U isometry, S Vh carries the weights
    assert np.max(np.abs(u.conj().T @ u - np.eye(u.shape[1]))) < 1e-12


def _run_heev(kern, dev, ks, mats):
    nb, kmax = len(ks), max(ks)
    sq = kmax * kmax
    ek = torch.tensor(ks, dtype=torch.int32, device=dev)
    g = torch.zeros(nb * sq, dtype=torch.complex128, device=dev)
    for t, (k, m) in enumerate(zip(ks, mats)):
        g[t * sq: t * sq + k * k] = torch.from_numpy(np.ascontiguousarray(m).reshape(-1)).to(dev)
    qh = torch.zeros(nb * sq, dtype=torch.complex128, device=dev)
    zt = torch.zeros(nb * sq, dtype=torch.complex128, device=dev)
    qa = torch.zeros(nb * sq, dtype=torch.complex128, device=dev)
    zr = torch.zeros(nb * sq, dtype=torch.float64, device=dev)
    d = torch.zeros(nb * kmax, dtype=torch.float64, device=dev)
    e = torch.zeros(nb * kmax, dtype=torch.float64, device=dev)
    tau = torch.zeros(nb * kmax, dtype=torch.complex128, device=dev)
    meta = torch.zeros(nb * 8, dtype=torch.int32, device=dev)
    status = torch.zeros(nb, dtype=torch.int32, device=dev)
    ext = Ext(ek, 0, 1)
    kern.heev(nb, None, ext, Buf(g, 0, sq), Buf(qh, 0, sq), Buf(zt, 0, sq), zr, sq, d, e, tau, kmax, meta, status, kmax)
    kern.gemm(K.OP_EVEC, nb, 1, None, ext, None, None, 1, 1, 1, Buf(zt, 0, sq), Buf(qh, 0, sq), Buf(qa, 0, sq), 0, kmax, kmax)
    return d.cpu().numpy().reshape(nb, kmax), qa.cpu().numpy().reshape(nb, sq), meta.cpu().numpy().reshape(nb, 8), status.cpu().numpy()


@pytest.mark.parametrize("ks", [[1, 2, 3, 5, 33], [64, 130, 17, 128], [256, 200], [512, 300]])
def test_heev_eigendecomposition(backends, ks):
    """Householder tridiagonalisation + implicit QL: rows of Qacc are the conjugated eigenvectors, d the eigenvalues."""
    cuda, _ = backends
    rng = np.random.default_rng(sum(ks))
    mats = []
    for i, k in enumerate(ks):
        r = max(1, k // 2) if i % 2 else k                 # every other matrix is a rank-deficient Gram matrix
        a = rng.standard_normal((k, r)) + 1j * rng.standard_normal((k, r))
        a *= np.logspace(0, -6, r)[None, :]                 # graded spectrum like a Schmidt spectrum
        mats.append(a @ a.conj().T)
    d, qa, meta, status = _run_heev(cuda, "cuda", ks, mats)
    assert not status.any()
    for t, k in enumerate(ks):
        assert meta[t, 5] == 1
        E = qa[t, : k * k].reshape(k, k).conj().T           # columns = eigenvectors
        lam = d[t, :k]
        scale = np.abs(np.linalg.eigvalsh(mats[t])).max()
        assert np.max(np.abs(E.conj().T @ E - np.eye(k))) < 1e-12 * max(1, k / 16)
        assert np.max(np.abs(mats[t] @ E - E * lam[None, :])) < 2e-13 * scale * max(1, k / 16)
        assert np.max(np.abs(np.sort(lam) - np.linalg.eigvalsh(mats[t]))) < 1e-13 * scale * max(1, k / 16)


def _run_heev_rows(kern, dev, ks, mats, mask=None):
    nb, kmax = len(ks), max(ks)
    sq = kmax * kmax
    ek = torch.tensor(ks, dtype=torch.int32, device=dev)
    g = torch.zeros(nb * sq, dtype=torch.complex128, device=dev)
    for t, (k, m) in enumerate(zip(ks, mats)):
        g[t * sq: t * sq + k * k] = torch.from_numpy(np.ascontiguousarray(m).reshape(-1)).to(dev)
    eh = torch.full((nb * sq,), 3.0 + 0j, dtype=torch.complex128, device=dev)
    zr = torch.zeros(nb * sq, dtype=torch.float64, device=dev)
    d = torch.zeros(nb * kmax, dtype=torch.float64, device=dev)
    e = torch.zeros(nb * kmax, dtype=torch.float64, device=dev)
    tau = torch.zeros(nb * kmax, dtype=torch.complex128, device=dev)
    meta = torch.zeros(nb * 8, dtype=torch.int32, device=dev)
    status = torch.zeros(nb, dtype=torch.int32, device=dev)
    mk = None if mask is None else torch.tensor(mask, dtype=torch.int32, device=dev)
    kern.heev_rows(nb, mk, Ext(ek, 0, 1), Buf(g, 0, sq), Buf(eh, 0, sq), zr, sq, d, e, tau, kmax, meta, status, kmax)
    return d.cpu().numpy().reshape(nb, kmax), eh.cpu().numpy().reshape(nb, sq), meta.cpu().numpy().reshape(nb, 8), status.cpu().numpy()


@pytest.mark.parametrize("ks", [[1, 2, 3, 5, 33], [16, 17, 31, 32, 34], [64, 130, 17, 128], [256, 200, 255], [512, 300], [257, 48]])
def test_heev_rows_eigendecomposition(backends, ks):
    """yb_heev_rows (blocked zlatrd panels + divide & conquer + reflectors applied to the rows of Z^T): row c of the output is
    (eigenvector c)^H.  Graded and rank-deficient Gram matrices as on the SVD route; sizes around every panel / tile boundary."""
    cuda, _ = backends
    rng = np.random.default_rng(sum(ks) + 1)
    mats = []
    for i, k in enumerate(ks):
        r = max(1, k // 2) if i % 2 else k
        a = rng.standard_normal((k, r)) + 1j * rng.standard_normal((k, r))
        a *= np.logspace(0, -6, r)[None, :]
        mats.append(a @ a.conj().T)
    mask = [1] * len(ks)
    d, eh, meta, status = _run_heev_rows(cuda, "cuda", ks, mats, mask)
    assert not status.any()
    for t, k in enumerate(ks):
        assert meta[t, 5] == 1
        E = eh[t, : k * k].reshape(k, k).conj().T           # columns = eigenvectors
        lam = d[t, :k]
        scale = np.abs(np.linalg.eigvalsh(mats[t])).max()
        assert np.max(np.abs(E.conj().T @ E - np.eye(k))) < 1e-12 * max(1, k / 16)
        assert np.max(np.abs(mats[t] @ E - E * lam[None, :])) < 2e-13 * scale * max(1, k / 16)
        assert np.max(np.abs(np.sort(lam) - np.linalg.eigvalsh(mats[t]))) < 1e-13 * scale * max(1, k / 16)


def test_heev_rows_special_matrices_and_mask(backends):
    """Diagonal, identity, 2 x 2 with a complex off-diagonal, a matrix with an exactly zero first column below the diagonal (tau = 0
    reflector); a masked trajectory is left untouched."""
    cuda, _ = backends
    rng = np.random.default_rng(5)
    k = 40
    diag = np.diag(np.linspace(1.0, 2.0, k)).astype(np.complex128)
    ident = np.eye(k, dtype=np.complex128)
    two = np.array([[2.0, 1.0 - 2.0j], [1.0 + 2.0j, -1.0]], dtype=np.complex128)
    a = rng.standard_normal((k, k)) + 1j * rng.standard_normal((k, k))
    blockd = a @ a.conj().T
    blockd[1:, 0] = 0
    blockd[0, 1:] = 0
    mats, ks = [diag, ident, two, blockd, blockd], [k, k, 2, k, k]
    d, eh, meta, status = _run_heev_rows(cuda, "cuda", ks, mats, [1, 1, 1, 1, 0])
    assert not status.any()
    for t, kk in enumerate(ks[:4]):
        E = eh[t, : kk * kk].reshape(kk, kk).conj().T
        lam = d[t, :kk]
        scale = max(1.0, np.abs(mats[t]).max())
        assert np.max(np.abs(E.conj().T @ E - np.eye(kk))) < 1e-12
        assert np.max(np.abs(mats[t] @ E - E * lam[None, :])) < 1e-12 * scale * kk
    assert np.all(eh[4] == 3.0)


@pytest.mark.parametrize("ks", [[256, 200, 131], [512, 97, 64, 33, 7]])
def test_heev_row_ring_is_bit_identical(backends, ks):
    """The cp.async row ring of k_hetrd / k_ungtr only changes HOW the trailing block is streamed: same operations in the same
    order, so eigenvalues and eigenvectors must be bit-identical to the direct-load variant."""
    cuda, _ = backends
    rng = np.random.default_rng(11 + sum(ks))
    mats = []
    for k in ks:
        a = rng.standard_normal((k, k)) + 1j * rng.standard_normal((k, k))
        a *= np.logspace(0, -7, k)[None, :]
        mats.append(a @ a.conj().T)
    try:
        cuda.set_option("heev_ring", 0)
        d0, qa0, meta0, status0 = _run_heev(cuda, "cuda", ks, mats)
        assert not status0.any()
        for depth in (6, 12, 16):
            cuda.set_option("heev_ring", depth)
            d1, qa1, meta1, status1 = _run_heev(cuda, "cuda", ks, mats)
            assert not status1.any()
            assert np.array_equal(d0, d1) and np.array_equal(qa0, qa1) and np.array_equal(meta0, meta1), f"ring depth {depth}"
    finally:
        cuda.set_option("heev_ring", -1)


def test_heev_diagonal_and_identity(backends):
    cuda, _ = backends
    ks = [4, 6, 40]
    mats = [np.diag([3.0, 1.0, 2.0, 0.0]).astype(np.complex128), np.eye(6, dtype=np.complex128), np.zeros((40, 40), dtype=np.complex128)]
    d, qa, meta, status = _run_heev(cuda, "cuda", ks, mats)
    assert not status.any()
    assert np.allclose(np.sort(d[0, :4]), [0, 1, 2, 3]) and np.allclose(d[1, :6], 1.0) and np.allclose(d[2, :40], 0.0)
    for t, k in enumerate(ks):
        E = qa[t, : k * k].reshape(k, k)
        assert np.max(np.abs(E @ E.conj().T - np.eye(k))) < 1e-13


@pytest.mark.parametrize("ks", [[65, 97, 160], [256, 255, 129]])
def test_heev_clustered_spectra(backends, ks):
    """Divide & conquer deflation paths: exactly degenerate clusters, a projector, a glued block-diagonal matrix."""
    cuda, _ = backends
    rng = np.random.default_rng(7 + sum(ks))
    mats = []
    for i, k in enumerate(ks):
        u, _ = np.linalg.qr(rng.standard_normal((k, k)) + 1j * rng.standard_normal((k, k)))
        if i == 0:
            lam = np.repeat([1.0, 0.5, 0.25, 1e-9, 0.0], k // 5 + 1)[:k]          # five exactly degenerate clusters
        elif i == 1:
            lam = (np.arange(k) < k // 3).astype(float)                            # projector
        else:
            lam = 1.0 + 1e-14 * rng.standard_normal(k)                             # one cluster at rounding distance
        mats.append((u * lam[None, :]) @ u.conj().T)
    d, qa, meta, status = _run_heev(cuda, "cuda", ks, mats)
    assert not status.any()
    for t, k in enumerate(ks):
        assert meta[t, 5] == 1
        E = qa[t, : k * k].reshape(k, k).conj().T
        lam = d[t, :k]
        assert np.max(np.abs(E.conj().T @ E - np.eye(k))) < 1e-12 * max(1, k / 16)
        assert np.max(np.abs(mats[t] @ E - E * lam[None, :])) < 2e-13 * max(1, k / 16)
        assert np.max(np.abs(np.sort(lam) - np.linalg.eigvalsh(mats[t]))) < 1e-13 * max(1, k / 16)


@pytest.mark.parametrize("dl,dr", [(1, 2), (4, 4), (2, 1), (1, 1)])
def test_mpo_site_apply_matches_twin(backends, dl, dr):
    """yb_mpo_site_apply (one site of MPO.multiply on an MPS, mpo_utils.py:27-56) on ragged bonds with a mask: bit-comparable sums of
    at most two products per element, and against the plain einsum of the definition."""
    cuda, twin = backends
    rng = np.random.default_rng(31 + dl * 7 + dr)
    nb, d, cap = 7, 2, 9
    ext = torch.from_numpy(rng.integers(1, cap + 1, size=(nb, 2)).astype(np.int32))
    mask = torch.tensor([1, 1, 0, 1, 1, 1, 1], dtype=torch.int32)
    slot_in, slot_out = d * cap * cap, d * cap * dl * cap * dr
    a = _rand_c(rng, nb * slot_in)
    w = _rand_c(rng, d * d * dl * dr)
    res = {}
    for name, kern, dev in (("cpu", twin, "cpu"), ("gpu", cuda, "cuda")):
        e, m = ext.to(dev), mask.to(dev)
        out = torch.zeros(nb * slot_out, dtype=torch.complex128, device=dev)
        kern.mpo_site_apply(nb, m, Ext(e, 0, 2), Ext(e, 1, 2), d, dl, dr, w.to(dev), Buf(a.clone().to(dev), 0, slot_in), Buf(out, 0, slot_out),
                            cap * dl * cap * dr)
        res[name] = out
    torch.cuda.synchronize()
    _cmp(res["cpu"], res["gpu"], tol=1e-14)
    got = res["gpu"].cpu().numpy()
    wt = w.numpy().reshape(d, d, dl, dr)
    for b in range(nb):
        l, r = int(ext[b, 0]), int(ext[b, 1])
        seg = got[b * slot_out: b * slot_out + d * l * dl * r * dr]
        if not mask[b]:
            assert not seg.any()
            continue
        want = np.einsum("stab,tlr->slarb", wt, a.numpy()[b * slot_in: b * slot_in + d * l * r].reshape(d, l, r)).reshape(-1)
        assert np.max(np.abs(seg - want)) < 1e-13


@pytest.mark.parametrize("case", GEMM_CASES)
@pytest.mark.parametrize("cap", [5, 37, 113])
@pytest.mark.parametrize("option", ["gemm_ragged", "gemm_warps16"])
def test_gemm_ragged_option_is_bit_identical(backends, case, cap, option):
    """Tiling options "gemm_ragged" (interleaved 8 x 8 sub-tiles, padding sub-tiles skipped) and "gemm_warps16" (16 warps with 16 x 32
    warp tiles instead of 8 with 32 x 32): every output element is the same sum in the same order, so the result must equal the
    default tiling bit for bit on ragged and full operands."""
    cuda, _ = backends
    op, ngroups, dphys, Dl, Dr, conj = case
    rng = np.random.default_rng(2000 * op + cap)
    nb = 5
    ext = torch.from_numpy(rng.integers(1, cap + 1, size=(nb, 3)).astype(np.int32))
    ext[0] = cap
    mask = torch.tensor([1, 1, 0, 1, 1], dtype=torch.int32).cuda()
    sa, sb, sc = _gemm_sizes(op, ngroups, dphys, Dl, Dr, cap)
    a, b = _rand_c(rng, nb * sa).cuda(), _rand_c(rng, nb * sb).cuda()
    c0 = _rand_c(rng, nb * sc)
    e = ext.cuda()
    maxm = 4 * 5 * cap
    outs = []
    try:
        for flag in (0, 1):
            cuda.set_option(option, flag)
            c = c0.clone().cuda()
            cuda.gemm(op, nb, ngroups, mask, Ext(e, 0, 3), Ext(e, 1, 3), Ext(e, 2, 3), dphys, Dl, Dr, Buf(a, 0, sa), Buf(b, 0, sb), Buf(c, 0, sc),
                      conj, maxm, maxm)
            torch.cuda.synchronize()
            outs.append(c.cpu().numpy())
    finally:
        cuda.set_option(option, 0)
    assert np.array_equal(outs[0], outs[1])


@pytest.mark.parametrize("dims", [(4, 4, 3, 3), (2, 2, 5, 4), (2, 2, 1, 1), (4, 4, 5, 5), (4, 4, 4, 3)])
def test_wapply_onepass_option_is_bit_identical(backends, dims):
    """A/B option "wapply_onepass": one read of every input element, all nx * ny outputs per thread; the same (u, v) order and zero skip
    as k_wapply, so the result must equal the default kernel bit for bit (ragged bonds, mask, sparse MPO tensor)."""
    cuda, _ = backends
    nx, nu, ny, nv = dims
    rng = np.random.default_rng(17)
    nb, cap = 6, 41
    ext = torch.from_numpy(rng.integers(1, cap + 1, size=(nb, 2)).astype(np.int32))
    ext[0] = cap
    mask = torch.tensor([1, 1, 0, 1, 1, 1], dtype=torch.int32).cuda()
    w = _rand_c(rng, nx * nu * ny * nv)
    w[::3] = 0
    sin, sout = nu * nv * cap * cap, nx * ny * cap * cap
    src, dst0 = _rand_c(rng, nb * sin).cuda(), _rand_c(rng, nb * sout)
    e = ext.cuda()
    outs = []
    try:
        for flag in (0, 1):
            cuda.set_option("wapply_onepass", flag)
            dst = dst0.clone().cuda()
            cuda.wapply(nb, mask, Ext(e, 0, 2), Ext(e, 1, 2), nx, nu, ny, nv, w.cuda(), Buf(src, 0, sin), Buf(dst, 0, sout), cap * cap)
            torch.cuda.synchronize()
            outs.append(dst.cpu().numpy())
    finally:
        cuda.set_option("wapply_onepass", 0)
    assert np.array_equal(outs[0], outs[1])


def test_workspace_bytes_covers_the_engine_allocations(backends):
    """What yb_workspace_bytes promises is what BatchedTJM allocates (within 2 %: a few small per-trajectory vectors)."""
    from yaqs_b200.engine import BatchedTJM, EngineParams
    from yaqs_b200.kernels import workspace_bytes
    cuda, _ = backends
    L, B, chi = 12, 6, 32
    mpo = [np.zeros((2, 2, 1 if i == 0 else 3, 1 if i == L - 1 else 3), dtype=np.complex128) for i in range(L)]
    init = [np.ones((2, 1, 1), dtype=np.complex128) for _ in range(L)]
    noise = [{"name": "pauli_z", "sites": [i], "strength": 0.1, "matrix": np.diag([1.0, -1.0]).astype(np.complex128)} for i in range(L)]
    torch.cuda.synchronize()
    before = torch.cuda.memory_allocated()
    eng = BatchedTJM(cuda, L, B, mpo, EngineParams(max_bond_dim=chi, order=2), noise, [(np.diag([1.0, -1.0]).astype(np.complex128), [0])], init)
    torch.cuda.synchronize()
    used = torch.cuda.memory_allocated() - before
    want = workspace_bytes(L, B, eng.lay.cap, eng.lay.D, 2, False)["total"]
    assert used <= 1.02 * want + (1 << 20) and want <= 1.02 * used + (1 << 20), (used, want)


def _cholqr_chain(kern, dev, mode, dist, dphys, ext_np, src_np, slot, thr, sd, mask_np=None):
    """prepare -> GRAM -> factor -> ROTATE -> GRAM -> refine -> ROTATE -> GRAM -> finalize on one backend (the engine's _svd_cholqr)."""
    nb = ext_np.shape[0]
    ext = torch.from_numpy(ext_np.copy()).to(dev)
    src = torch.from_numpy(src_np.copy()).to(dev)
    mask = None if mask_np is None else torch.from_numpy(mask_np.astype(np.int32)).to(dev)
    sq = sd * sd
    bufs = {n: torch.zeros(nb * sq, dtype=torch.complex128, device=dev) for n in ("w", "linv", "g", "y1", "g2", "y2", "u", "v")}
    meta = torch.zeros(nb * K.SVD_META, dtype=torch.int32, device=dev)
    rot = torch.zeros(nb, dtype=torch.int32, device=dev)
    flag = torch.full((nb,), 7, dtype=torch.int32, device=dev)
    nflag = torch.zeros(1, dtype=torch.int32, device=dev)
    keep = torch.full((nb,), -1, dtype=torch.int32, device=dev)
    B = lambda n: Buf(bufs[n], 0, sq)      # noqa: E731
    orient = K.SVD_ORIENT_COLS if dist == "right" else K.SVD_ORIENT_ROWS
    kern.svd_prepare(nb, mask, mode | orient, Ext(ext, 0, 2), Ext(ext, 1, 2), dphys, Buf(src, 0, slot), B("w"), B("linv"), meta, rot)
    ek, eln = Ext(meta, 0, K.SVD_META), Ext(meta, 1, K.SVD_META)
    kern.gemm(K.OP_GRAM, nb, 1, mask, ek, eln, None, 1, 1, 1, B("w"), B("w"), B("g"), 0, sd, sd)
    kern.cholqr_factor(nb, mask, meta, B("g"), B("linv"), thr, flag, nflag, min(sd, 128))
    kern.gemm(K.OP_ROTATE, nb, 1, mask, ek, eln, None, 1, 1, 1, B("linv"), B("w"), B("y1"), 0, sd, sd)
    kern.gemm(K.OP_GRAM, nb, 1, mask, ek, eln, None, 1, 1, 1, B("y1"), B("y1"), B("g2"), 0, sd, sd)
    kern.cholqr_refine(nb, mask, meta, B("g2"), B("g"), 1e-7, flag, nflag)
    kern.gemm(K.OP_ROTATE, nb, 1, mask, ek, eln, None, 1, 1, 1, B("g"), B("y1"), B("y2"), 0, sd, sd)
    kern.gemm(K.OP_GRAM, nb, 1, mask, ek, eln, None, 1, 1, 1, B("w"), B("y2"), B("g2"), 0, sd, sd)
    kern.cholqr_finalize(nb, mask, meta, B("y2"), B("g2"), Ext(keep, 0, 1), B("u"), B("v"), flag, sd * 2 * sd)
    if dev != "cpu":
        torch.cuda.synchronize()
    return (flag.cpu().numpy(), int(nflag.item()), keep.cpu().numpy(), bufs["u"].cpu().numpy().reshape(nb, sq), bufs["v"].cpu().numpy().reshape(nb, sq),
            bufs["linv"].cpu().numpy().reshape(nb, sq), meta.cpu().numpy().reshape(nb, K.SVD_META))


@pytest.mark.parametrize("mode,dist", [(K.SVD_ROWS, "right"), (K.SVD_COLS, "left")])
def test_cholqr_gauge_shift_matches_twin_and_is_exact(backends, mode, dist):
    """Keep-all gauge shifts by Cholesky-QR (csrc/yb_cholqr.cuh): flags / certificate agree with the twin; for accepted trajectories
    keep = bond side, the isometric factor is orthonormal to 1e-14 and the two factors reproduce the centre tensor to 1e-13; graded
    spectra down to 1e-4 relative pass, rank-deficient and too-thin matrices are flagged, masked trajectories are untouched."""
    cuda, twin = backends
    rng = np.random.default_rng(5)
    d = 2
    # (l, r) per trajectory; bond side = r for ROWS / right, l for COLS / left
    shapes = [(64, 128), (128, 128), (20, 33), (1, 2), (7, 5), (100, 128), (64, 128), (3, 128), (64, 128), (37, 70)]
    kinds = ["random", "graded", "random", "random", "random", "rankdef", "tiny_sv", "thin", "masked", "graded"]
    if mode == K.SVD_COLS:
        shapes = [(r, l) for l, r in shapes]
    nb = len(shapes)
    sd = 256
    slot = sd * sd
    ext = np.array(shapes, dtype=np.int32)
    src = np.zeros(nb * slot, dtype=np.complex128)
    mats = []
    for t, ((l, r), kind) in enumerate(zip(shapes, kinds)):
        rows, cols = (d * l, r) if mode == K.SVD_ROWS else (l, d * r)
        kb = cols if mode == K.SVD_ROWS else rows
        M = rng.standard_normal((rows, cols)) + 1j * rng.standard_normal((rows, cols))
        if kind in ("graded", "rankdef", "tiny_sv") and min(rows, cols) == kb:
            u, s, vh = np.linalg.svd(M, full_matrices=False)
            s = 10.0 ** (-np.linspace(0, 4, kb))
            if kind == "rankdef":
                s[-3:] = 0.0
            if kind == "tiny_sv":
                s[-1] = 3e-7                                  # sigma^2 = 9e-14 < 1e-12: must be truncated, i.e. NOT certified
            M = (u * s) @ vh
        M /= np.linalg.norm(M)
        mats.append(M)
        if mode == K.SVD_ROWS:
            src[t * slot: t * slot + M.size] = M.reshape(-1)
        else:                                                 # l x (s r) stored as (s, l, r)
            src[t * slot: t * slot + M.size] = M.reshape(l, d, r).transpose(1, 0, 2).reshape(-1)
    mask = np.array([0 if k == "masked" else 1 for k in kinds])
    fg, ng, kg, ug, vg, lg, mg = _cholqr_chain(cuda, "cuda:0", mode, dist, d, ext, src, slot, 1e-12, sd, mask)
    fc, nc, kc, uc, vc, lc, mc = _cholqr_chain(twin, "cpu", mode, dist, d, ext, src, slot, 1e-12, sd, mask)
    assert np.array_equal(fg, fc) and ng == nc, (fg, fc)
    for t, ((l, r), kind) in enumerate(zip(shapes, kinds)):
        M = mats[t]
        rows, cols = M.shape
        kb = cols if mode == K.SVD_ROWS else rows
        if kind == "masked":
            assert fg[t] == 0 and kg[t] == -1 and not ug[t].any()
            continue
        expect_flag = kind in ("rankdef", "tiny_sv", "thin") or kb > min(rows, cols)
        assert bool(fg[t]) == expect_flag, (t, kind, fg[t])
        assert mg[t, 7] == 0
        if expect_flag:
            assert kg[t] == -1
            continue
        assert kg[t] == kb
        U, V = ug[t, : rows * kb].reshape(rows, kb), vg[t, : kb * cols].reshape(kb, cols)
        iso = U.conj().T @ U if dist == "right" else V @ V.conj().T
        assert np.max(np.abs(iso - np.eye(kb))) < 1e-14, (t, kind, float(np.max(np.abs(iso - np.eye(kb)))))
        assert np.max(np.abs(U @ V - M)) < 1e-13, (t, kind)
        # the first-pass triangular inverse against LAPACK
        G = (M.conj().T @ M) if mode == K.SVD_ROWS else (M @ M.conj().T)
        Li = np.linalg.inv(np.linalg.cholesky(G))
        got = lg[t, : kb * kb].reshape(kb, kb)
        assert np.max(np.abs(got @ np.linalg.cholesky(G) - np.eye(kb))) < 1e-7, (t, kind)      # graded spectra: cond(L) ~ 1e4
        assert not np.triu(got, 1).any()


@pytest.mark.parametrize("ks,top", [([1, 2, 3, 5, 33], 4), ([16, 17, 31, 32, 34], 16), ([64, 130, 17, 128], 64), ([256, 200, 255], 128), ([512, 300], 256)])
def test_heev_rows_top_and_capped_finalize(backends, ks, top):
    """yb_heev_rows_top -> ROTATE_TOP -> yb_svd_finalize_top (the capped split of the Gram route): the compact block holds the
    eigenvector rows of the `top` largest eigenvalues in descending order; a capped truncation of W = random k x len matrices through
    that block gives LAPACK's singular values, its keep count, an isometric U and U (S Vh) = the best rank-keep approximation."""
    cuda, _ = backends
    dev = "cuda"
    rng = np.random.default_rng(sum(ks) + top)
    nb, kmax = len(ks), max(ks)
    sq = kmax * kmax
    mats = []
    for i, k in enumerate(ks):
        a = rng.standard_normal((k, k)) + 1j * rng.standard_normal((k, k))
        u, _, vh = np.linalg.svd(a)
        sv = np.logspace(0, -5 if i % 2 else -2, k)
        mats.append((u * sv) @ vh)
    ek = torch.tensor(ks, dtype=torch.int32, device=dev)
    w = torch.zeros(nb * sq, dtype=torch.complex128, device=dev)
    meta = torch.zeros(nb * 8, dtype=torch.int32, device=dev)
    for t, (k, m) in enumerate(zip(ks, mats)):
        w[t * sq: t * sq + k * k] = torch.from_numpy(np.ascontiguousarray(m).reshape(-1)).to(dev)
        meta[t * 8: t * 8 + 8] = torch.tensor([k, k, 0, k, k, 0, 0, 0], dtype=torch.int32)        # W = M (orient 0), rows = cols = k
    g = torch.zeros(nb * sq, dtype=torch.complex128, device=dev)
    eh = torch.full((nb * sq,), 3.0 + 0j, dtype=torch.complex128, device=dev)
    wf = torch.zeros(nb * sq, dtype=torch.complex128, device=dev)
    zr = torch.zeros(nb * sq, dtype=torch.float64, device=dev)
    d = torch.zeros(nb * kmax, dtype=torch.float64, device=dev)
    e = torch.zeros(nb * kmax, dtype=torch.float64, device=dev)
    ds = torch.zeros(nb * kmax, dtype=torch.float64, device=dev)
    order = torch.zeros(nb * kmax, dtype=torch.int32, device=dev)
    ntop = torch.zeros(nb, dtype=torch.int32, device=dev)
    tau = torch.zeros(nb * kmax, dtype=torch.complex128, device=dev)
    status = torch.zeros(nb, dtype=torch.int32, device=dev)
    keep = torch.zeros(nb, dtype=torch.int32, device=dev)
    sval = torch.zeros(nb * kmax, dtype=torch.float64, device=dev)
    uo = torch.zeros(nb * sq, dtype=torch.complex128, device=dev)
    vo = torch.zeros(nb * sq, dtype=torch.complex128, device=dev)
    X = Ext(ek, 0, 1)
    cuda.gemm(K.OP_GRAM, nb, 1, None, X, X, None, 1, 1, 1, Buf(w, 0, sq), Buf(w, 0, sq), Buf(g, 0, sq), 0, kmax, kmax)
    cuda.heev_rows_top(nb, None, X, Buf(g, 0, sq), Buf(eh, 0, sq), zr, sq, d, e, tau, kmax, meta, status, kmax, top, order, ds, ntop)
    cuda.gemm(K.OP_ROTATE_TOP, nb, 1, None, X, X, Ext(ntop, 0, 1), 1, 1, 1, Buf(eh, 0, sq), Buf(w, 0, sq), Buf(wf, 0, sq), 0, min(kmax, top), kmax)
    thr = 1e-7
    cuda.svd_finalize_top(nb, None, meta, 0, thr, top, 1, 1, Buf(wf, 0, sq), Buf(eh, 0, sq), sval, kmax, Ext(keep, 0, 1), Buf(uo, 0, sq), Buf(vo, 0, sq),
                          status, kmax, ds, kmax, top)
    torch.cuda.synchronize()
    assert not status.cpu().numpy().any()
    from oracle.tjm_oracle import truncate
    for t, k in enumerate(ks):
        m = mats[t]
        nt = min(k, top)
        assert int(ntop[t]) == nt
        G = m @ m.conj().T
        lam = np.linalg.eigvalsh(G)[::-1]
        scale = lam[0]
        assert np.max(np.abs(ds.cpu().numpy()[t * kmax: t * kmax + k] - lam)) < 1e-13 * scale * max(1, k / 16)
        E = eh.cpu().numpy()[t * sq: t * sq + nt * k].reshape(nt, k).conj().T       # columns = top eigenvectors, descending
        assert np.max(np.abs(E.conj().T @ E - np.eye(nt))) < 1e-12 * max(1, k / 16)
        assert np.max(np.abs(G @ E - E * lam[None, :nt])) < 2e-13 * scale * max(1, k / 16)
        u_ref, s_ref, vh_ref = np.linalg.svd(m)
        kp = truncate(s_ref, "discarded_weight", thr, top, 1)
        assert int(keep[t]) == kp, (t, k, int(keep[t]), kp)
        got_s = sval.cpu().numpy()[t * kmax: t * kmax + k]
        assert np.max(np.abs(got_s[:nt] - s_ref[:nt])) < 1e-12
        assert nt == k or np.max(np.abs(got_s[nt:] ** 2 - s_ref[nt:] ** 2)) < 1e-12
        U = uo.cpu().numpy()[t * sq: t * sq + k * kp].reshape(k, kp)
        V = vo.cpu().numpy()[t * sq: t * sq + kp * k].reshape(kp, k)
        assert np.max(np.abs(U.conj().T @ U - np.eye(kp))) < 1e-12 * max(1, k / 16)
        best = (u_ref[:, :kp] * s_ref[:kp]) @ vh_ref[:kp]
        assert np.max(np.abs(U @ V - best)) < 1e-10


def test_jump_test_and_select_match_twin(backends):
    """yb_jump_test (table and in-kernel Philox modes: dp, the uniforms, the mask, stream positions, the jump count) and yb_jump_select
    (weights from the local density matrices, pdf, numpy's Generator.choice rule, gathered operator rows, zero-weight status) against
    the numpy twin; the Philox mode also against the table yb_philox_uniform fills."""
    cuda, twin = backends
    rng = np.random.default_rng(11)
    B, L, nproc, nops = 37, 6, 13, 24
    norm2 = torch.from_numpy(rng.uniform(0.5, 1.0, B))
    table = torch.from_numpy(rng.random((B, 40)))
    pos0 = torch.from_numpy(rng.integers(0, 30, B).astype(np.int64))
    traj = torch.from_numpy(rng.integers(0, 2 ** 40, B).astype(np.int64))

    def run_test(kern, dev, mode):
        nd = norm2.to(dev)
        out = {n: torch.zeros(B, dtype=torch.float64, device=dev) for n in ("dp", "u1", "u2")}
        jm, nj = torch.zeros(B, dtype=torch.int32, device=dev), torch.full((1,), -5, dtype=torch.int32, device=dev)
        pos = pos0.clone().to(dev)
        if mode == "table_pos":
            kern.jump_test(B, nd, table.to(dev), 0, pos, None, out["dp"], out["u1"], out["u2"], jm, nj)
        elif mode == "table_col":
            kern.jump_test(B, nd, table.to(dev), 6, None, None, out["dp"], out["u1"], out["u2"], jm, nj)
        elif mode == "philox_pos":
            kern.jump_test(B, nd, None, 0, pos, (77, traj.to(dev), 0x5452414A, 0), out["dp"], out["u1"], out["u2"], jm, nj)
        else:
            kern.jump_test(B, nd, None, 0, None, (77, traj.to(dev), 0x53414D50, 9), out["dp"], out["u1"], out["u2"], jm, nj)
        return {k: v.cpu().numpy() for k, v in out.items()}, jm.cpu().numpy(), int(nj.item()), pos.cpu().numpy()

    for mode in ("table_pos", "table_col", "philox_pos", "philox_col"):
        og, jg, ng, pg = run_test(cuda, "cuda", mode)
        oc, jc, nc, pc = run_test(twin, "cpu", mode)
        for k in og:
            assert np.array_equal(og[k], oc[k]), (mode, k)
        assert np.array_equal(jg, jc) and ng == nc == int(jg.sum()) and np.array_equal(pg, pc), mode
        if mode.endswith("pos"):
            assert np.array_equal(pg, pos0.numpy() + 1 + jg)
        assert np.array_equal(jg.astype(bool), og["u1"] < og["dp"])
    # Philox mode == the table yb_philox_uniform fills (positions across block boundaries included)
    tab = torch.zeros((B, 40), dtype=torch.float64, device="cuda")
    cuda.philox_uniform(77, traj.cuda(), 0x5452414A, 20, 0, tab)
    og, _, _, _ = run_test(cuda, "cuda", "philox_pos")
    tn = tab.cpu().numpy()
    assert np.array_equal(og["u1"], tn[np.arange(B), pos0.numpy()]) and np.array_equal(og["u2"], tn[np.arange(B), pos0.numpy() + 1])

    # ---- selection
    rdm1 = torch.from_numpy(rng.standard_normal((L, B, 4)) + 1j * rng.standard_normal((L, B, 4)))
    rdm1[..., 0] = rdm1[..., 0].abs() + 0.1
    rdm1[..., 3] = rdm1[..., 3].abs() + 0.1
    rdm2 = torch.from_numpy(rng.standard_normal((L - 1, B, 16)) + 1j * rng.standard_normal((L - 1, B, 16)))
    psite = torch.from_numpy(rng.integers(0, L - 1, nproc).astype(np.int32))
    ptwo = torch.from_numpy((rng.random(nproc) < 0.3).astype(np.int32))
    ldl = np.zeros((nproc, 16), dtype=np.complex128)
    for k in range(nproc):
        m = rng.standard_normal((4, 4)) + 1j * rng.standard_normal((4, 4)) if ptwo[k] else np.pad(np.diag(rng.uniform(0.2, 1.0, 2)), ((0, 2), (0, 2)))
        mm = m.conj().T @ m if ptwo[k] else m
        ldl[k] = mm.reshape(-1) if ptwo[k] else np.concatenate([mm[:2, :2].reshape(-1), np.zeros(12)])
    pldl = torch.from_numpy(ldl)
    for k in range(nproc):                                    # keep two-site weights positive: rho_2 = identity-like
        if ptwo[k]:
            rdm2[psite[k]] = torch.from_numpy(np.tile(np.eye(4).reshape(-1), (B, 1)) * rng.uniform(0.5, 1.0, (B, 1)))
    strength = torch.from_numpy(rng.uniform(0.05, 0.3, nproc))
    ops = torch.from_numpy(rng.standard_normal((nproc, nops)) + 1j * rng.standard_normal((nproc, nops)))
    jmask = torch.from_numpy((rng.random(B) < 0.6).astype(np.int32))
    jmask[0] = 1
    u2 = torch.from_numpy(rng.random(B))
    u2[0] = 0.999999999
    zero_w = 3                                                # one jumping trajectory with all weights zero -> status bit 4
    jmask[zero_w] = 1
    rdm1[:, zero_w] = 0
    rdm2[:, zero_w] = 0

    def run_sel(kern, dev):
        w = torch.zeros((B, nproc), dtype=torch.float64, device=dev)
        tot = torch.zeros(B, dtype=torch.float64, device=dev)
        ch = torch.zeros(B, dtype=torch.int32, device=dev)
        sel = torch.zeros(B * nops, dtype=torch.complex128, device=dev)
        status = torch.zeros(B, dtype=torch.int32, device=dev)
        kern.jump_select(B, jmask.to(dev), u2.to(dev), rdm1.contiguous().to(dev), rdm2.contiguous().to(dev), nproc, psite.to(dev), ptwo.to(dev), pldl.to(dev),
                         strength.to(dev), 0.1, ops.contiguous().to(dev), nops, w, tot, ch, sel, status)
        return w.cpu().numpy(), tot.cpu().numpy(), ch.cpu().numpy(), sel.cpu().numpy().reshape(B, nops), status.cpu().numpy()

    wg, tg, cg, sg, stg = run_sel(cuda, "cuda")
    wc, tc, cc, sc, stc = run_sel(twin, "cpu")
    jm = jmask.numpy().astype(bool)
    assert np.array_equal(cg, cc) and np.array_equal(stg, stc) and stg[zero_w] == 4 and not stg[np.arange(B) != zero_w].any()
    ok = jm & (np.arange(B) != zero_w)
    assert np.max(np.abs(tg[ok] - tc[ok])) < 1e-13 and np.max(np.abs(wg[ok] - wc[ok])) < 1e-13
    assert np.array_equal(sg[ok], ops.numpy()[cg[ok]]) and (cg[~jm] == -1).all()
    for b in np.nonzero(ok)[0]:                                # numpy's own choice rule on the kernel's pdf
        cdf = np.cumsum(wg[b]); cdf /= cdf[-1]
        assert cg[b] == min(int(np.searchsorted(cdf, u2[b].item(), side="right")), nproc - 1)
