"""numpy twin of the C ABI in include/yaqs_b200.h -- TEST INFRASTRUCTURE ONLY.

Same entry points, same buffer layouts, same per-trajectory extents as ``yaqs_b200.kernels.CudaKernels``, evaluated
with plain numpy loops over the batch on CPU torch tensors.  It exists so that
  * ``-m gpu`` tests can check every CUDA kernel against an independent statement of its contract, and
  * ``-m "not gpu"`` tests can run the host-side engine (sweep order, masks, gauge bookkeeping) against the golden
    fixtures without a GPU.
It is never imported by the product package; ``yaqs_b200`` fails loudly when the CUDA library is missing.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg
import torch

from oracle.tjm_oracle import truncate as oracle_truncate

LANCZOS_PITCH = 32
SVD_META = 8
_TRUNC_NAMES = {0: "discarded_weight", 1: "relative", 2: "hard_cutoff", 3: "relative_discarded_weight"}


def _np(t: torch.Tensor) -> np.ndarray:
    return t.view(-1).numpy()


def _ext(e, b):
    return int(_np(e.tensor)[e.offset + b * e.stride])


def _view(buf, b, n):
    a = _np(buf.tensor)
    start = buf.offset + b * buf.stride
    return a[start:start + n]


def _on(mask, b):
    return mask is None or int(_np(mask)[b]) != 0


def philox4x32_10(ctr: np.ndarray, key: np.ndarray) -> np.ndarray:
    """Vectorised Philox4x32-10 (Salmon et al., SC'11): ctr [..., 4] uint32, key [..., 2] uint32 -> [..., 4] uint32."""
    c = [ctr[..., i].astype(np.uint64) for i in range(4)]
    k0, k1 = key[..., 0].astype(np.uint64), key[..., 1].astype(np.uint64)
    m32 = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = np.uint64(0xD2511F53) * c[0], np.uint64(0xCD9E8D57) * c[2]
        c = [(p1 >> np.uint64(32)) ^ c[1] ^ k0, p1 & m32, (p0 >> np.uint64(32)) ^ c[3] ^ k1, p0 & m32]
        k0, k1 = (k0 + np.uint64(0x9E3779B9)) & m32, (k1 + np.uint64(0xBB67AE85)) & m32
    return np.stack(c, axis=-1).astype(np.uint32)


def philox_pairs(seed: int, traj: int, tag: int, blocks: np.ndarray) -> np.ndarray:
    """[len(blocks), 2] float64: the two uniforms of every block of stream (seed, traj, tag)."""
    n = len(blocks)
    ctr = np.empty((n, 4), dtype=np.uint32)
    ctr[:, 0] = np.asarray(blocks, dtype=np.uint64) & np.uint64(0xFFFFFFFF)
    ctr[:, 1] = tag & 0xFFFFFFFF
    t = traj & (2 ** 64 - 1)
    ctr[:, 2], ctr[:, 3] = t & 0xFFFFFFFF, t >> 32
    s = seed & (2 ** 64 - 1)
    key = np.broadcast_to(np.array([s & 0xFFFFFFFF, s >> 32], dtype=np.uint32), (n, 2))
    r = philox4x32_10(ctr, key).astype(np.uint64)
    hi = np.stack([r[:, 0], r[:, 2]], axis=1) >> np.uint64(5)
    lo = np.stack([r[:, 1], r[:, 3]], axis=1) >> np.uint64(6)
    return ((hi << np.uint64(26)) | lo).astype(np.float64) / 9007199254740992.0


class NumpyKernels:
    name = "numpy-twin"

    def __init__(self, device=None):
        self.device = torch.device("cpu")
        self._launches = 0
        self._svd_store: dict[int, tuple] = {}

    def launch_count(self, reset=False):
        v = self._launches
        if reset:
            self._launches = 0
        return v

    def sync_int(self, t):
        return int(t.item())

    # -- contractions -------------------------------------------------------------------------------------------------
    def gemm(self, op, nbatch, ngroups, mask, e0, e1, e2, dphys, Dl, Dr, a, b, c, conj_b, max_m, max_n):
        self._launches += 1
        for t in range(nbatch):
            if not _on(mask, t):
                continue
            x0 = _ext(e0, t) if e0 is not None else 0
            x1 = _ext(e1, t) if e1 is not None else 0
            x2 = _ext(e2, t) if e2 is not None else 0
            cj = (lambda z: z.conj()) if conj_b else (lambda z: z)
            if op == 0:      # RCONTRACT: X[(p,l),r] R[r,(a,r')]
                X = _view(a, t, dphys * x0 * x1).reshape(dphys * x0, x1)
                R = cj(_view(b, t, x1 * Dr * x1).reshape(x1, Dr * x1))
                _view(c, t, dphys * x0 * Dr * x1)[:] = (X @ R).reshape(-1)
            elif op == 1:    # LCONTRACT: Y_o[l',r'] = sum_(l,a) Lenv[(l,a),l'] T2_o[(l,a),r']
                Lm = _view(a, t, x0 * Dl * x0).reshape(x0 * Dl, x0)
                T2 = cj(_view(b, t, dphys * x0 * Dl * x1).reshape(dphys, x0 * Dl, x1))
                _view(c, t, dphys * x0 * x1)[:] = np.einsum("km,okn->omn", Lm, T2).reshape(-1)
            elif op == 2:    # ENVR3: R'[(l,a),l'] = sum_(o,r') T2_o[(l,a),r'] conj(bra_o[l',r'])
                T2 = _view(a, t, dphys * x0 * Dl * x1).reshape(dphys, x0 * Dl, x1)
                bra = _view(b, t, dphys * x0 * x1).reshape(dphys, x0, x1)
                bra = bra if conj_b else bra.conj()
                _view(c, t, x0 * Dl * x0)[:] = np.einsum("omk,onk->mn", T2, bra).reshape(-1)
            elif op == 3:    # ENVL1: U1_o[(l,a),r'] = sum_l' Lenv[(l,a),l'] conj(bra_o[l',r'])
                Lm = _view(a, t, x0 * Dl * x0).reshape(x0 * Dl, x0)
                bra = _view(b, t, dphys * x0 * x1).reshape(dphys, x0, x1)
                bra = bra if conj_b else bra.conj()
                _view(c, t, dphys * x0 * Dl * x1)[:] = np.einsum("mk,okn->omn", Lm, bra).reshape(-1)
            elif op == 4:    # ENVL3: L'[r,(a',r')] = sum_(p,l) ket[(p,l),r] U2[(p,l),(a',r')]
                ket = _view(a, t, dphys * x0 * x1).reshape(dphys * x0, x1)
                U2 = cj(_view(b, t, dphys * x0 * Dr * x1).reshape(dphys * x0, Dr * x1))
                _view(c, t, x1 * Dr * x1)[:] = (ket.T @ U2).reshape(-1)
            elif op == 5:    # MERGE
                A = _view(a, t, dphys * x0 * x1).reshape(dphys, x0, x1)
                Bm = cj(_view(b, t, dphys * x1 * x2).reshape(dphys, x1, x2))
                _view(c, t, dphys * dphys * x0 * x2)[:] = np.einsum("slc,tcr->stlr", A, Bm).reshape(-1)
            elif op == 6:    # ABSORB_R: next'_s[k,r] = G[k,j] next_s[j,r]
                G = _view(a, t, x0 * x1).reshape(x0, x1)
                nx = cj(_view(b, t, ngroups * x1 * x2).reshape(ngroups, x1, x2))
                _view(c, t, ngroups * x0 * x2)[:] = np.einsum("kj,sjr->skr", G, nx).reshape(-1)
            elif op == 7:    # ABSORB_L: prev'_s[l,k] = prev_s[l,j] G[j,k]
                pv = _view(a, t, ngroups * x0 * x1).reshape(ngroups, x0, x1)
                G = cj(_view(b, t, x1 * x2).reshape(x1, x2))
                _view(c, t, ngroups * x0 * x2)[:] = np.einsum("slj,jk->slk", pv, G).reshape(-1)
            elif op == 8:    # GRAM: G = W W^H
                W = _view(a, t, x0 * x1).reshape(x0, x1)
                W2 = _view(b, t, x0 * x1).reshape(x0, x1)
                _view(c, t, x0 * x0)[:] = (W @ W2.conj().T).reshape(-1)
            elif op == 9:    # EVEC: Qacc[c,j] = sum_i Zt[c,i] conj(Qh[j,i])
                Zt = _view(a, t, x0 * x0).reshape(x0, x0)
                Qh = _view(b, t, x0 * x0).reshape(x0, x0)
                _view(c, t, x0 * x0)[:] = (Zt @ Qh.conj().T).reshape(-1)
            elif op == 10:   # ROTATE: Wf = Qacc W
                Qa = _view(a, t, x0 * x0).reshape(x0, x0)
                W = _view(b, t, x0 * x1).reshape(x0, x1)
                _view(c, t, x0 * x1)[:] = (Qa @ W).reshape(-1)
            elif op == 11:   # ROTATE_TOP: Wf[c] = Etop[c] W for the x2 compact eigenvector rows
                nr = min(x2, x0)
                Qa = _view(a, t, nr * x0).reshape(nr, x0)
                W = _view(b, t, x0 * x1).reshape(x0, x1)
                _view(c, t, nr * x1)[:] = (Qa @ W).reshape(-1)
            else:
                raise ValueError(op)

    def rcontract_mpo(self, nbatch, mask, el, er, dphys, d_left, d_right, x, renv, wt, out, max_l, max_r):
        """Twin of yb_rcontract_mpo: out[x,l,y,r'] = sum_(u,v) wt[x,u,y,v] sum_r X[u,l,r] R[r,v,r']."""
        self._launches += 1
        W = _np(wt)[: dphys * dphys * d_left * d_right].reshape(dphys, dphys, d_left, d_right)
        for t in range(nbatch):
            if not _on(mask, t):
                continue
            l, r = _ext(el, t), _ext(er, t)
            X = _view(x, t, dphys * l * r).reshape(dphys * l, r)
            R = _view(renv, t, r * d_right * r).reshape(r, d_right * r)
            T1 = (X @ R).reshape(dphys, l, d_right, r)
            _view(out, t, dphys * l * d_left * r)[:] = np.einsum("xuyv,ulvr->xlyr", W, T1).reshape(-1)

    def wapply(self, nbatch, mask, el, er, nx, nu, ny, nv, wt, src, dst, max_lr):
        self._launches += 1
        W = _np(wt)[: nx * nu * ny * nv].reshape(nx, nu, ny, nv)
        for t in range(nbatch):
            if not _on(mask, t):
                continue
            l, r = _ext(el, t), _ext(er, t)
            X = _view(src, t, nu * l * nv * r).reshape(nu, l, nv, r)
            _view(dst, t, nx * l * ny * r)[:] = np.einsum("xuyv,ulvr->xlyr", W, X).reshape(-1)

    def env_identity(self, nbatch, mask, echi, dmpo, env, max_chi):
        self._launches += 1
        for t in range(nbatch):
            if not _on(mask, t):
                continue
            chi = _ext(echi, t)
            e = np.zeros((chi, dmpo, chi), dtype=np.complex128)
            for i in range(chi):
                e[i, :, i] = 1
            _view(env, t, e.size)[:] = e.reshape(-1)

    # -- Lanczos ------------------------------------------------------------------------------------------------------
    def lanczos_init(self, nbatch, mask, el, er, dphys, x, basis, stride_vec, nrm, active, kdone, n_zero=None):
        self._launches += 1
        if n_zero is not None:
            _np(n_zero)[0] = 0
        for t in range(nbatch):
            if not _on(mask, t):
                _np(active)[t] = 0
                _np(kdone)[t] = -1
                continue
            n = dphys * _ext(el, t) * _ext(er, t)
            v = _view(x, t, n)
            nr = float(np.linalg.norm(v))
            _np(nrm)[t] = nr
            _np(active)[t] = 1 if nr > 0 else 0
            _np(kdone)[t] = 0
            if nr > 0:
                _view(basis, t, n)[:] = v / nr

    def lanczos_step(self, nbatch, j, kmax, dt, tol, el, er, dphys, basis, stride_vec, alpha, beta, nrm, active, kdone, coef, n_active, n_zero=None):
        self._launches += 1
        if n_zero is not None:
            _np(n_zero)[0] = 0
        al = _np(alpha).reshape(-1, LANCZOS_PITCH)
        be = _np(beta).reshape(-1, LANCZOS_PITCH)
        cf = _np(coef).reshape(-1, LANCZOS_PITCH)
        for t in range(nbatch):
            if not int(_np(active)[t]):
                continue
            n = dphys * _ext(el, t) * _ext(er, t)
            vj = _view(basis.at(j * stride_vec), t, n)
            w = _view(basis.at((j + 1) * stride_vec), t, n)
            aj = float(np.vdot(vj, w).real)
            al[t, j] = aj
            w -= aj * vj
            if j > 0:
                w -= be[t, j - 1] * _view(basis.at((j - 1) * stride_vec), t, n)
            last = j == kmax - 1
            bj = float(np.linalg.norm(w))
            if not last:
                be[t, j] = bj
            k = j + 1
            breakdown = (not last) and bj < 100.0 * n * np.finfo(np.float64).eps
            finish = False
            if breakdown or j >= 1 or last:
                if k == 1:
                    lam, z = np.array([aj]), np.array([[1.0]])
                else:
                    lam, z = scipy.linalg.eigh_tridiagonal(al[t, :k], be[t, :k - 1])
                ph = np.exp(-1j * dt * lam) * z[0]
                if breakdown or last:
                    finish = True
                elif bj * abs(np.dot(z[k - 1], ph)) < tol:
                    finish = True
                if finish:
                    cf[t, :k] = float(_np(nrm)[t]) * (z @ ph)
            if finish:
                _np(kdone)[t] = k
                _np(active)[t] = 0
            else:
                w /= bj
                _np(n_active)[0] += 1

    def lanczos_combine(self, nbatch, mask, el, er, dphys, basis, stride_vec, coef, kmax, kdone, out):
        self._launches += 1
        cf = _np(coef).reshape(-1, LANCZOS_PITCH)
        for t in range(nbatch):
            if not _on(mask, t):
                continue
            k = int(_np(kdone)[t])
            if k <= 0:
                continue
            n = dphys * _ext(el, t) * _ext(er, t)
            acc = np.zeros(n, dtype=np.complex128)
            for i in range(k):
                acc += cf[t, i] * _view(basis.at(i * stride_vec), t, n)
            _view(out, t, n)[:] = acc

    # -- SVD (LAPACK stands in for the Jacobi rounds; outputs are gauge-equivalent) ------------------------------------
    def svd_prepare(self, nbatch, mask, mode, el, er, dphys, src, w, q, meta, rot_count, n_zero=None, status_zero=None):
        self._launches += 1
        if n_zero is not None:
            _np(n_zero)[0] = 0
        if status_zero is not None:
            _np(status_zero)[:nbatch] = 0
        mt = _np(meta).reshape(-1, SVD_META)
        force, mode = (mode >> 4) & 3, mode & 3
        for t in range(nbatch):
            if not _on(mask, t):
                mt[t, 5] = 1
                continue
            l, r = _ext(el, t), _ext(er, t)
            if mode == 0:
                M = _view(src, t, dphys * dphys * l * r).reshape(dphys, dphys, l, r).transpose(0, 2, 1, 3).reshape(dphys * l, dphys * r)
            elif mode == 1:
                M = _view(src, t, dphys * l * r).reshape(dphys * l, r)
            else:
                M = _view(src, t, dphys * l * r).reshape(dphys, l, r).transpose(1, 0, 2).reshape(l, dphys * r)
            rows, cols = M.shape
            orient = 0 if rows <= cols else 1
            if force:
                orient = force - 1
            k, ln = (cols, rows) if orient else (rows, cols)
            mt[t] = [k, ln, orient, rows, cols, 0, 0, 0]
            W = M.conj().T if orient else M
            _view(w, t, k * ln)[:] = np.ascontiguousarray(W).reshape(-1)
            _view(q, t, k * k)[:] = np.eye(k, dtype=np.complex128).reshape(-1)
            _np(rot_count)[t] = 0

    def svd_round(self, nbatch, mask, meta, block_rows, nblocks, rnd, full_panel, inner_sweeps, w, q, rot_count, max_len, max_k, n_zero=None):
        self._launches += 1          # the twin factorises in svd_finalize with LAPACK
        if n_zero is not None:
            _np(n_zero)[0] = 0

    def svd_check(self, nbatch, mask, meta, rot_count, n_pending):
        self._launches += 1
        mt = _np(meta).reshape(-1, SVD_META)
        for t in range(nbatch):
            if _on(mask, t) and not mt[t, 5]:
                mt[t, 5] = 1
                mt[t, 6] += 1

    def svd_finalize(self, nbatch, mask, meta, trunc_mode, threshold, max_bond, min_keep, dist, w, q, sval, stride_s, keep, uout, vout, status, max_k):
        self._launches += 1
        mt = _np(meta).reshape(-1, SVD_META)
        for t in range(nbatch):
            if not _on(mask, t):
                continue
            k, ln, orient, rows, cols = (int(v) for v in mt[t, :5])
            W = _view(w, t, k * ln).reshape(k, ln)
            Qm = _view(q, t, k * k).reshape(k, k)          # identity on the Jacobi route (the twin's rounds are no-ops)
            W = Qm.conj().T @ W
            M = W.conj().T if orient else W
            u, s, vh = scipy.linalg.svd(M, full_matrices=False, lapack_driver="gesdd", check_finite=False)
            kp = oracle_truncate(s, _TRUNC_NAMES[trunc_mode], threshold, max_bond if max_bond > 0 else None, min_keep)
            _np(sval)[t * stride_s: t * stride_s + k] = 0.0
            _np(sval)[t * stride_s: t * stride_s + len(s)] = s
            _np(keep.tensor)[keep.offset + t * keep.stride] = kp
            u, sk, vh = u[:, :kp], s[:kp], vh[:kp]
            if dist == 0:
                u = u * sk
            elif dist == 1:
                vh = vh * sk[:, None]
            else:
                u = u * np.sqrt(sk)
                vh = vh * np.sqrt(sk)[:, None]
            _view(uout, t, rows * kp)[:] = np.ascontiguousarray(u).reshape(-1)
            _view(vout, t, kp * cols)[:] = np.ascontiguousarray(vh).reshape(-1)

    def svd_ambiguous(self, nbatch, mask, meta, sval, stride_s, trunc_mode, threshold, two_level, tau_rel, noise_rel, status, amb, n_amb):
        """Twin of k_svd_ambiguous (csrc/yb_svd.cu): same comparisons, same suffix-sum order."""
        self._launches += 1
        mt = _np(meta).reshape(-1, SVD_META)
        count = 0
        for t in range(nbatch):
            if not _on(mask, t):
                _np(amb)[t] = 0
                continue
            nsv = int(min(mt[t, 3], mt[t, 4]))
            s2 = _np(sval)[t * stride_s: t * stride_s + nsv] ** 2
            smax2 = float(s2[0]) if nsv else 0.0
            err_hi = noise_rel * smax2
            err = np.where(s2 >= tau_rel * smax2, err_hi, err_hi * tau_rel if two_level else err_hi)
            flag = (mt[t, 5] == 0) or (status is not None and _np(status)[t] != 0) or not np.isfinite(smax2)
            if trunc_mode == 0:
                if threshold > 0.0 and nsv:
                    cum = np.cumsum(s2[::-1])[::-1]
                    flag = flag or bool(np.any(np.abs(cum - threshold) <= err + 1e-9 * threshold))
            elif nsv:
                flag = flag or bool(np.any(np.abs(s2 - threshold * threshold * smax2) <= err))
            _np(amb)[t] = 1 if flag else 0
            count += int(bool(flag))
        _np(n_amb)[0] += count

    def svd_select(self, nbatch, mask, meta, w, tau_rel, sel_n, sel_idx, max_k):
        self._launches += 1
        mt = _np(meta).reshape(-1, SVD_META)
        for t in range(nbatch):
            if not _on(mask, t):
                _np(sel_n)[t] = 0
                continue
            k, ln = int(mt[t, 0]), int(mt[t, 1])
            nu = (np.abs(_view(w, t, k * ln).reshape(k, ln)) ** 2).sum(axis=1)
            idx = np.nonzero(nu < tau_rel * nu.max())[0] if k else np.zeros(0, dtype=int)
            _np(sel_n)[t] = len(idx) if len(idx) >= 2 else 0
            _np(sel_idx)[t * max_k: t * max_k + len(idx)] = idx

    def svd_gather(self, nbatch, mask, meta, sel_n, sel_idx, max_k, direction, w, q, ws, qs):
        self._launches += 1
        mt = _np(meta).reshape(-1, SVD_META)
        for t in range(nbatch):
            n = int(_np(sel_n)[t])
            if not _on(mask, t) or n <= 0:
                continue
            k, ln = int(mt[t, 0]), int(mt[t, 1])
            idx = _np(sel_idx)[t * max_k: t * max_k + n]
            W, Q = _view(w, t, k * ln).reshape(k, ln), _view(q, t, k * k).reshape(k, k)
            Ws, Qs = _view(ws, t, n * ln).reshape(n, ln), _view(qs, t, n * k).reshape(n, k)
            if direction == 0:
                Ws[:], Qs[:] = W[idx], Q[idx]
            else:
                W[idx], Q[idx] = Ws, Qs

    def heev(self, nbatch, mask, ek, g, qh, zt, zr, stride_zr, d, e, tau, stride_vec, meta, status, max_k):
        """LAPACK stands in for tridiagonalisation + QL: Qh = I, row c of zt = conj-free eigenvector c (so EVEC gives E^H)."""
        self._launches += 3
        mt = _np(meta).reshape(-1, SVD_META) if meta is not None else None
        for t in range(nbatch):
            if not _on(mask, t):
                continue
            k = _ext(ek, t)
            G = _view(g, t, k * k).reshape(k, k)
            lam, vec = np.linalg.eigh(0.5 * (G + G.conj().T))
            _view(qh, t, k * k)[:] = np.eye(k, dtype=np.complex128).reshape(-1)
            _view(zt, t, k * k)[:] = np.ascontiguousarray(vec.conj().T).reshape(-1)
            _np(d)[t * stride_vec: t * stride_vec + k] = lam
            if mt is not None:
                mt[t, 5] = 1

    def heev_rows(self, nbatch, mask, ek, g, eh, zr, stride_zr, d, e, tau, stride_vec, meta, status, max_k):
        """Twin of yb_heev_rows: LAPACK eigh; row c of eh = (eigenvector c)^H."""
        self._launches += 3
        mt = _np(meta).reshape(-1, SVD_META) if meta is not None else None
        for t in range(nbatch):
            if not _on(mask, t):
                continue
            k = _ext(ek, t)
            G = _view(g, t, k * k).reshape(k, k)
            lam, vec = np.linalg.eigh(0.5 * (G + G.conj().T))
            _view(eh, t, k * k)[:] = np.ascontiguousarray(vec.conj().T).reshape(-1)
            _np(d)[t * stride_vec: t * stride_vec + k] = lam
            if mt is not None:
                mt[t, 5] = 1

    def heev_rows_top(self, nbatch, mask, ek, g, eh, zr, stride_zr, d, e, tau, stride_vec, meta, status, max_k, top, order, dsorted, ntop):
        """Twin of yb_heev_rows_top: eigenvector rows of the `top` largest eigenvalues (compact, descending), the descending spectrum."""
        self._launches += 4
        mt = _np(meta).reshape(-1, SVD_META) if meta is not None else None
        for t in range(nbatch):
            if not _on(mask, t):
                continue
            k = _ext(ek, t)
            G = _view(g, t, k * k).reshape(k, k)
            lam, vec = np.linalg.eigh(0.5 * (G + G.conj().T))
            idx = np.argsort(-lam, kind="stable")
            nt = min(k, top)
            _view(eh, t, nt * k)[:] = np.ascontiguousarray(vec[:, idx[:nt]].conj().T).reshape(-1)
            _np(d)[t * stride_vec: t * stride_vec + k] = lam
            _np(order)[t * stride_vec: t * stride_vec + k] = idx
            _np(dsorted)[t * stride_vec: t * stride_vec + k] = lam[idx]
            _np(ntop)[t] = nt
            if mt is not None:
                mt[t, 5] = 1

    def svd_finalize_top(self, nbatch, mask, meta, trunc_mode, threshold, max_bond, min_keep, dist, w, q, sval, stride_s, keep, uout, vout, status, max_k,
                         lam, stride_lam, top):
        """Twin of yb_svd_finalize_top: the kept factors from the compact top block (w = Etop W, q = Etop), the singular values behind
        it from the Gram eigenvalues."""
        self._launches += 1
        mt = _np(meta).reshape(-1, SVD_META)
        for t in range(nbatch):
            if not _on(mask, t):
                continue
            k, ln, orient, rows, cols = (int(v) for v in mt[t, :5])
            nt = min(k, top)
            Wf = _view(w, t, nt * ln).reshape(nt, ln)
            Q = _view(q, t, nt * k).reshape(nt, k)
            sig = np.sqrt((np.abs(Wf) ** 2).sum(axis=1))
            perm = np.argsort(-sig, kind="stable")
            tail = np.sqrt(np.maximum(_np(lam)[t * stride_lam + nt: t * stride_lam + k], 0.0))
            s = np.concatenate([sig[perm], tail])
            nsv = min(k, rows, cols)
            kp = oracle_truncate(s[:nsv], _TRUNC_NAMES[trunc_mode], threshold, max_bond if max_bond > 0 else None, min_keep)
            assert kp <= nt
            _np(sval)[t * stride_s: t * stride_s + k] = s
            _np(keep.tensor)[keep.offset + t * keep.stride] = kp
            sk = s[:kp]
            f = {0: sk, 1: np.ones(kp), 2: np.sqrt(sk)}[dist]            # factor on U
            gq = {0: np.ones(kp), 1: sk, 2: np.sqrt(sk)}[dist]           # factor on Vh
            with np.errstate(all="ignore"):
                inv = np.where(sk > 0, 1.0 / np.where(sk > 0, sk, 1.0), 0.0)
            Qk, Wk = Q[perm[:kp]], Wf[perm[:kp]]
            if orient == 0:      # W = M: U[:, c] = conj(Q[c]) f, Vh[c] = Wf[c] g / sigma
                u = Qk.conj().T * f
                vh = Wk * (gq * inv)[:, None]
            else:                # W = M^H: U[:, c] = conj(Wf[c]) f / sigma, Vh[c] = Q[c] g
                u = Wk.conj().T * (f * inv)
                vh = Qk * gq[:, None]
            _view(uout, t, rows * kp)[:] = np.ascontiguousarray(u).reshape(-1)
            _view(vout, t, kp * cols)[:] = np.ascontiguousarray(vh).reshape(-1)

    # -- truncation-free gauge shifts by Cholesky-QR (twin of csrc/yb_cholqr.cuh: same certificate, same outputs) ---------
    def cholqr_factor(self, nbatch, mask, meta, g, linv, threshold, flag, n_flag, max_k):
        self._launches += 1
        mt = _np(meta).reshape(-1, SVD_META)
        for t in range(nbatch):
            if not _on(mask, t):
                _np(flag)[t] = 0
                continue
            k, ln = int(mt[t, 0]), int(mt[t, 1])
            ok = 1 <= k <= min(ln, max_k)
            if ok:
                G = _view(g, t, k * k).reshape(k, k)
                H = np.tril(G) + np.tril(G, -1).conj().T                  # the kernel reads the lower triangle only
                H[np.diag_indices(k)] = G.diagonal().real
                trace = float(H.diagonal().real.sum())
                try:
                    with np.errstate(all="ignore"):
                        L = np.linalg.cholesky(H)
                        Li = scipy.linalg.solve_triangular(L, np.eye(k, dtype=np.complex128), lower=True, check_finite=False)
                    fn = float((np.abs(Li) ** 2).sum())
                    lb = 1.0 / fn - 64.0 * k * 2.220446049250313e-16 * trace
                    ok = bool(np.isfinite(fn) and np.isfinite(trace) and trace > 0.0 and lb >= 4.0 * threshold and lb > 0.0)
                    if ok:
                        _view(linv, t, k * k)[:] = np.ascontiguousarray(Li).reshape(-1)
                except np.linalg.LinAlgError:
                    ok = False
            _np(flag)[t] = 0 if ok else 1
            if not ok:
                _np(n_flag)[0] += 1

    def cholqr_refine(self, nbatch, mask, meta, g2, x, tol, flag, n_flag):
        self._launches += 1
        mt = _np(meta).reshape(-1, SVD_META)
        for t in range(nbatch):
            if not _on(mask, t) or _np(flag)[t]:
                continue
            k = int(mt[t, 0])
            G2 = _view(g2, t, k * k).reshape(k, k)
            with np.errstate(all="ignore"):
                e2 = float((np.abs(G2 - np.eye(k)) ** 2).sum())
            _view(x, t, k * k)[:] = (1.5 * np.eye(k) - 0.5 * G2).reshape(-1)
            if not e2 <= tol * tol:
                _np(flag)[t] = 1
                _np(n_flag)[0] += 1

    def cholqr_finalize(self, nbatch, mask, meta, y2, l2, keep, uout, vout, flag, max_elems):
        self._launches += 1
        mt = _np(meta).reshape(-1, SVD_META)
        for t in range(nbatch):
            if not _on(mask, t):
                continue
            mt[t, 7] = 0
            if _np(flag)[t]:
                continue
            k, ln, orient = int(mt[t, 0]), int(mt[t, 1]), int(mt[t, 2])
            mt[t, 5] = 1
            Y2 = _view(y2, t, k * ln).reshape(k, ln)
            L2 = _view(l2, t, k * k).reshape(k, k)
            _np(keep.tensor)[keep.offset + t * keep.stride] = k
            if orient:
                _view(uout, t, ln * k)[:] = np.ascontiguousarray(Y2.conj().T).reshape(-1)
                _view(vout, t, k * k)[:] = np.ascontiguousarray(L2.conj().T).reshape(-1)
            else:
                _view(uout, t, k * k)[:] = np.ascontiguousarray(L2).reshape(-1)
                _view(vout, t, k * ln)[:] = np.ascontiguousarray(Y2).reshape(-1)

    # -- site-tensor utilities ----------------------------------------------------------------------------------------
    def unpack_vh(self, nbatch, mask, ek, er, dphys, src, dst, max_elems, set_ext=None):
        self._launches += 1
        for t in range(nbatch):
            if not _on(mask, t):
                continue
            k, r = _ext(ek, t), _ext(er, t)
            v = _view(src, t, k * dphys * r).reshape(k, dphys, r)
            _view(dst, t, k * dphys * r)[:] = np.ascontiguousarray(v.transpose(1, 0, 2)).reshape(-1)
            if set_ext is not None:
                _np(set_ext.tensor)[set_ext.offset + t * set_ext.stride] = k

    def copy_packed(self, nbatch, mask, e0, e1, mult, src, dst, max_elems, set_ext=None):
        self._launches += 1
        for t in range(nbatch):
            if not _on(mask, t):
                continue
            n = mult * _ext(e0, t) * (_ext(e1, t) if e1 is not None else 1)
            _view(dst, t, n)[:] = _view(src, t, n)
            if set_ext is not None:
                _np(set_ext.tensor)[set_ext.offset + t * set_ext.stride] = _ext(e0, t)

    def apply_local_op(self, nbatch, mask, sel, sel_value, el, er, dphys, op, op_stride, a, max_lr):
        self._launches += 1
        ops = _np(op)
        for t in range(nbatch):
            if not _on(mask, t):
                continue
            if sel is not None and int(_np(sel)[t]) != sel_value:
                continue
            l, r = _ext(el, t), _ext(er, t)
            o = ops[t * op_stride: t * op_stride + dphys * dphys].reshape(dphys, dphys)
            v = _view(a, t, dphys * l * r).reshape(dphys, l * r)
            v[:] = o @ v

    def mpo_site_apply(self, nbatch, mask, el, er, dphys, d_left, d_right, w, src, dst, max_out_lr):
        self._launches += 1
        wt = _np(w)[: dphys * dphys * d_left * d_right].reshape(dphys, dphys, d_left, d_right)
        for t in range(nbatch):
            if not _on(mask, t):
                continue
            l, r = _ext(el, t), _ext(er, t)
            a = _view(src, t, dphys * l * r).reshape(dphys, l, r)
            out = np.einsum("stab,tlr->slarb", wt, a).reshape(-1)
            _view(dst, t, out.size)[:] = out

    def norm2(self, nbatch, mask, el, er, dphys, a, out):
        self._launches += 1
        for t in range(nbatch):
            if not _on(mask, t):
                continue
            v = _view(a, t, dphys * _ext(el, t) * _ext(er, t))
            _np(out)[t] = float(np.vdot(v, v).real)

    def scale(self, nbatch, mask, el, er, dphys, scale, a, max_elems, rsqrt_mode=False):
        self._launches += 1
        for t in range(nbatch):
            if not _on(mask, t):
                continue
            v = _view(a, t, dphys * _ext(el, t) * _ext(er, t))
            f = float(_np(scale)[t])
            with np.errstate(all="ignore"):
                v *= (1.0 / np.sqrt(f)) if rsqrt_mode else f

    # -- stochastic_process on the device (twin of k_jump_test / k_jump_select) -------------------------------------------
    def jump_test(self, nbatch, norm2, table, col, pos, philox, dp, u1, u2, jmask, njump):
        self._launches += 1
        count = 0
        for t in range(nbatch):
            p = int(_np(pos)[t]) if pos is not None else int(col)
            if philox is not None:
                seed, traj, tag, block0 = philox
                q = 2 * int(block0) + p
                pr = philox_pairs(int(seed), int(_np(traj)[t]), int(tag), np.array([q >> 1, (q >> 1) + 1], dtype=np.uint64)).reshape(-1)
                a, b = (pr[1], pr[2]) if (q & 1) else (pr[0], pr[1])
            else:
                row = table[t].numpy() if table.dim() == 2 else table.numpy()
                a, b = float(row[p]), float(row[p + 1])
            d = 1.0 - float(_np(norm2)[t])
            j = 1 if a < d else 0
            _np(dp)[t], _np(u1)[t], _np(u2)[t], _np(jmask)[t] = d, a, b, j
            if pos is not None:
                _np(pos)[t] = p + 1 + j
            count += j
        _np(njump)[0] = count

    def jump_select(self, nbatch, jmask, u2, rdm1, rdm2, nproc, proc_site, proc_two, proc_ldl, proc_strength, dt, ops_table, nops, w, total, choice,
                    ops_sel, status):
        self._launches += 1
        site, two = _np(proc_site), _np(proc_two)
        ldl = _np(proc_ldl).reshape(nproc, 16)
        strength = _np(proc_strength)
        g1 = _np(rdm1).reshape(-1, nbatch, 4)
        g2 = _np(rdm2).reshape(-1, nbatch, 16) if rdm2 is not None else None
        wv = _np(w).reshape(-1, nproc)
        ops = _np(ops_table).reshape(nproc, nops)
        sel = _np(ops_sel).reshape(-1, nops)
        for t in range(nbatch):
            if not _np(jmask)[t]:
                _np(total)[t], _np(choice)[t] = 0.0, -1
                continue
            tot = 0.0
            for k in range(nproc):
                if two[k]:
                    acc = float(np.sum((ldl[k] * g2[site[k], t]).real))
                else:
                    acc = float(np.sum((ldl[k, :4] * g1[site[k], t]).real))
                wv[t, k] = acc * (dt * float(strength[k]))
                tot += wv[t, k]
            _np(total)[t] = tot
            ok = np.isfinite(tot) and tot > 0.0
            if not ok:
                _np(status)[t] |= 4
            with np.errstate(all="ignore"):
                wv[t] = wv[t] / (tot if ok else 1.0)
                cdf = np.cumsum(wv[t])
                ch = int(np.sum((cdf / cdf[-1]) <= float(_np(u2)[t])))
            ch = min(ch, nproc - 1)
            _np(choice)[t] = ch
            sel[t] = ops[ch]

    def observables(self, nbatch, nobs, col, site, two, mats, rdm1, rdm2, out, out_stride, status):
        self._launches += 1
        m = _np(mats).reshape(nobs, 16)
        g1 = _np(rdm1).reshape(-1, nbatch, 4)
        g2 = _np(rdm2).reshape(-1, nbatch, 16) if rdm2 is not None else None
        o = _np(out)
        for k in range(nobs):
            s_, c = int(_np(site)[k]), int(_np(col)[k])
            for t in range(nbatch):
                v = complex(np.sum(m[k] * g2[s_, t])) if _np(two)[k] else complex(np.sum(m[k, :4] * g1[s_, t]))
                o[t * out_stride + c] = v.real
                if v.imag >= 1e-13:
                    _np(status)[t] |= 8

    def diagnostics(self, nbatch, length, dphys, bond, out, out_stride):
        self._launches += 1
        bd = _np(bond).reshape(-1, length + 1)[:nbatch, 1:length].astype(np.float64)
        o = _np(out)
        for t in range(nbatch):
            o[t * out_stride + 0] = float((bd[t] ** 3).sum())
            o[t * out_stride + 1] = float(max(dphys, bd[t].max())) if bd.shape[1] else float(dphys)
            o[t * out_stride + 2] = float(bd[t].sum())

    def mask_andnot(self, nbatch, mask, flag, work):
        self._launches += 1
        for t in range(nbatch):
            _np(work)[t] = 1 if (_on(mask, t) and not _np(flag)[t]) else 0

    def philox_uniform(self, seed, traj_index, tag, nblocks, block0, out):
        """numpy restatement of csrc/yb_philox.cuh (Philox4x32-10, key = seed, counter = (block, tag, traj_lo, traj_hi))."""
        o = out.numpy()
        for b, t in enumerate(_np(traj_index).tolist()):
            o[b, : 2 * nblocks] = philox_pairs(int(seed), int(t), int(tag), np.arange(block0, block0 + nblocks, dtype=np.uint64)).reshape(-1)

    def local_rdm(self, nbatch, mask, el, er, dphys, a, x, g):
        self._launches += 1
        gg = _np(g).reshape(-1, dphys, dphys)
        for t in range(nbatch):
            if not _on(mask, t):
                continue
            lr = _ext(el, t) * _ext(er, t)
            av = _view(a, t, dphys * lr).reshape(dphys, lr)
            xv = _view(x, t, dphys * lr).reshape(dphys, lr)
            gg[t] = av.conj() @ xv.T
