// Batched divide & conquer eigen-solver for real symmetric tridiagonal matrices (one CTA per trajectory).
//
// Role on the path: third stage of the Gram-matrix SVD (yb_heev.cu) that replaces scipy.linalg.svd (zgesdd) at
// core/linalg/svd.py:51-104 for the matrices of split_two_site (core/methods/decompositions.py:105-185) and the gauge shifts
// (core/data_structures/mps.py:747-764).  The implicit-QL kernel (k_steqr) applies ~k^2 dependent Givens rotations to a
// k x k matrix that does not fit on chip: 14.9 of the 31.5 ms of a 296 x (256 x 256) eigen-solve.  Cuppen's divide & conquer
// does the same job with O(k^2) scalar work per level and ONE dense product per merge:
//     T = diag(T1', T2') + rho u u^T          (cut at every multiple of 32: leaves of 32 solved by QL held in registers)
//     D + rho z z^T,  z = Q^T u               sort, deflate (tiny z_j; close poles: one Givens rotation), secular equation
//                                             (one thread per root: "middle way" rational iteration, bracketed), Gu/Eisenstat
//                                             recomputation of z, eigenvectors v_i[j] = zhat_j / (d_j - lam_i)
//     Q_new = Q_old[:, kept] V                64 x 128 register-tiled FP64 product, V generated on the fly while staging
// All per-merge vectors live in shared memory; the eigenvector matrices ping-pong between two k x k real global buffers.
// Storage is TRANSPOSED throughout: row c of Q holds eigenvector c, so every pass is a coalesced row operation.
//
// The body is written as barrier-separated phases over `tid` so that the SAME source compiles for the host with the
// threads of a phase run one after the other (YB_HOST_EMU, tests/test_stedc_host_emulation.py): the numerics of this
// kernel are checked against LAPACK on the CPU, the GPU tests check the parallel execution.
#pragma once
#include <math.h>

#ifdef YB_HOST_EMU
#include <algorithm>
#include <vector>
#define YB_DC_FN static inline
#define YB_THREADS_BEGIN for (int tid = 0; tid < nthr; ++tid) {
#define YB_THREADS_END }
#define YB_DC_ACC(r) acc_store[(size_t)tid * 32 + (r)]
#else
#define YB_DC_FN __device__ __forceinline__
#define YB_THREADS_BEGIN { const int tid = threadIdx.x;
#define YB_THREADS_END } __syncthreads();
#define YB_DC_ACC(r) acc[r]
#endif

namespace stedc {

constexpr int LEAF = 32;
constexpr int THREADS = 256;
constexpr int TM = 64, TN = 128, TK = 16;           // product tile: 64 new eigenvectors x 128 components, 16-deep chunks
constexpr double EPS = 2.220446049250313e-16;

template <int KM>
struct Smem {
    double sD[KM], se[KM], Ds[KM], zs[KM], rc[KM], rs[KM], dk[KM], zk[KM], mu[KM], zh[KM], vn[KM];
    int perm[KM], keepl[KM], defll[KM], ra[KM], rb[KM], org[KM], krow[KM];
    double As[TK][TM], Bs[TK][TN];
    double red[64];
    int info[8];          // 0: kept (K'), 1: deflated, 2: rotations, 3: fail, 4: all-deflated shortcut
};

// ---- leaf: implicit QL on a block of at most LEAF rows; this thread owns component `lane` of every eigenvector ----------
// (every lane runs the scalar recurrence redundantly -- identical values, no communication; z[] is indexed statically)
YB_DC_FN int leaf_ql(double* dl, double* el, double* z, int n, int lane) {
#pragma unroll
    for (int i = 0; i < LEAF; ++i) z[i] = (i == lane) ? 1.0 : 0.0;
    int l = 0, iter = 0;
    while (l < n) {
        int m = l;
        for (; m < n - 1; ++m) {
            const double dd = fabs(dl[m]) + fabs(dl[m + 1]);
            if (fabs(el[m]) <= EPS * dd) break;
        }
        if (m == l) { ++l; iter = 0; continue; }
        if (++iter > 60) return 1;
        double g = (dl[l + 1] - dl[l]) / (2.0 * el[l]);
        double r = sqrt(g * g + 1.0);
        g = dl[m] - dl[l] + el[l] / (g + copysign(r, g));
        double s = 1.0, c = 1.0, p = 0.0;
        bool under = false;
#pragma unroll
        for (int i = LEAF - 2; i >= 0; --i) {
            if (i < m && i >= l && !under) {
                double f = s * el[i];
                const double b = c * el[i];
                r = sqrt(f * f + g * g);
                el[i + 1] = r;
                if (r == 0.0) {
                    dl[i + 1] -= p;
                    el[m] = 0.0;
                    under = true;
                } else {
                    s = f / r;
                    c = g / r;
                    g = dl[i + 1] - p;
                    r = (dl[i] - g) * s + 2.0 * c * b;
                    p = s * r;
                    dl[i + 1] = g + p;
                    g = c * r - b;
                    f = z[i + 1];
                    z[i + 1] = s * z[i] + c * f;
                    z[i] = c * z[i] - s * f;
                }
            }
        }
        if (!under) { dl[l] -= p; el[l] = g; el[m] = 0.0; }
    }
    return 0;
}

// ---- secular equation: root i of 1 + rho sum_j z_j^2 / (d_j - lam) in (d_i, d_i+1); lam = d[org] + mu ------------------
YB_DC_FN int secular_root(int i, int K, const double* dk, const double* zk, double rho, int& org_out, double& mu_out) {
    if (K == 1) { org_out = 0; mu_out = rho * zk[0] * zk[0]; return 0; }
    int org, p0, p1;
    double lo, hi, mu;
    if (i < K - 1) {
        const double di = dk[i], half = 0.5 * (dk[i + 1] - di);
        double f = 0.0;
        for (int j = 0; j < K; ++j) f += zk[j] * zk[j] / ((dk[j] - di) - half);
        f = 1.0 + rho * f;
        if (f > 0.0) { org = i; lo = 0.0; hi = half; mu = half; }
        else { org = i + 1; lo = -half; hi = 0.0; mu = -half; }
        p0 = i; p1 = i + 1;
    } else {
        double zz = 0.0;
        for (int j = 0; j < K; ++j) zz += zk[j] * zk[j];
        org = K - 1; lo = 0.0; hi = rho * zz; mu = hi;
        p0 = K - 1; p1 = K;
    }
    const double dorg = dk[org];
    int rc = 1;
    for (int it = 0; it < 100; ++it) {
        double psi = 0.0, dpsi = 0.0, phi = 0.0, dphi = 0.0;
        for (int j = 0; j <= p0; ++j) {
            const double r = 1.0 / ((dk[j] - dorg) - mu), t = zk[j] * zk[j] * r;
            psi += t; dpsi += t * r;
        }
        for (int j = p0 + 1; j < K; ++j) {
            const double r = 1.0 / ((dk[j] - dorg) - mu), t = zk[j] * zk[j] * r;
            phi += t; dphi += t * r;
        }
        psi *= rho; dpsi *= rho; phi *= rho; dphi *= rho;
        const double g = 1.0 + psi + phi;
        const double err = 8.0 * EPS * (1.0 + fabs(psi) + fabs(phi)) + EPS * fabs(mu) * (dpsi + dphi);
        if (fabs(g) <= err) { rc = 0; break; }
        if (!(g == g)) break;                                   // NaN: give up (caller flags the trajectory)
        if (g > 0.0) hi = fmin(hi, mu); else lo = fmax(lo, mu);
        if (hi - lo <= 2.0 * EPS * fmax(fabs(lo), fabs(hi))) {
            mu = (lo != 0.0 && hi != 0.0) ? 0.5 * (lo + hi) : (lo == 0.0 ? hi : lo);
            rc = 0;
            break;
        }
        // "middle way": psi ~ s + S / (D0 - eta), phi ~ r + R / (D1 - eta) around the current point (eta = step in mu)
        const double D0 = (dk[p0] - dorg) - mu;
        const double S = dpsi * D0 * D0, s = psi - dpsi * D0;
        double eta;
        if (p1 < K) {
            const double D1 = (dk[p1] - dorg) - mu;
            const double R = dphi * D1 * D1, r = phi - dphi * D1;
            const double c = 1.0 + s + r;
            const double a = c, b = -(c * (D0 + D1) + S + R), cc = c * D0 * D1 + S * D1 + R * D0;
            if (a == 0.0) {
                eta = (b != 0.0) ? -cc / b : NAN;
            } else {
                double disc = b * b - 4.0 * a * cc;
                if (disc < 0.0) disc = 0.0;
                const double sq = sqrt(disc);
                const double q = (b <= 0.0) ? -0.5 * (b - sq) : -0.5 * (b + sq);
                const double r1 = q / a, r2 = (q != 0.0) ? cc / q : NAN;
                eta = (r1 > D0 && r1 < D1) ? r1 : r2;
            }
        } else {
            const double c = 1.0 + s + phi;
            eta = (c != 0.0) ? D0 + S / c : NAN;
        }
        double nw = mu + eta;
        if (!(nw > lo && nw < hi)) nw = 0.5 * (lo + hi);        // outside the bracket (or NaN): bisect
        mu = nw;
    }
    if (mu == 0.0) rc = 1;                                      // a root on a pole cannot happen after deflation
    org_out = org;
    mu_out = mu;
    return rc;
}

// ---- the whole solver for one matrix ------------------------------------------------------------------------------------
// d_io/e_in: tridiagonal (k, k-1 entries); QA, QB: two k*k real work matrices; on exit d_io = eigenvalues (unsorted) and the
// returned pointer (QA or QB) holds the eigenvectors, row c = eigenvector c.  Returns nullptr-equivalent flag via *fail_out.
template <int KM>
YB_DC_FN double* solve(Smem<KM>& S, double* d_io, const double* e_in, int k, double* QA, double* QB, int nthr, int* fail_out
#ifdef YB_HOST_EMU
                       , std::vector<double>& acc_store
#endif
) {
    // ---- load, scale to max-norm 1 ----
    YB_THREADS_BEGIN
    if (tid < 64) {
        double m = 0.0;
        for (int i = tid; i < k; i += 64) {
            const double a = fabs(d_io[i]), b = (i < k - 1) ? fabs(e_in[i]) : 0.0;
            m = fmax(m, (a == a) ? a : INFINITY);
            m = fmax(m, (b == b) ? b : INFINITY);
        }
        S.red[tid] = m;
    }
    if (tid == 0) { S.info[3] = 0; }
    YB_THREADS_END
    YB_THREADS_BEGIN
    if (tid == 0) {
        double m = 0.0;
        for (int i = 0; i < 64; ++i) m = fmax(m, S.red[i]);
        S.red[0] = m;
    }
    YB_THREADS_END
    const double anorm = S.red[0];
    if (!(anorm > 0.0) || !(anorm < INFINITY)) {
        // zero matrix: eigenvalues stay as they are (all zero), eigenvectors = identity; non-finite input: flag it
        YB_THREADS_BEGIN
        for (long long x = tid; x < (long long)k * k; x += nthr) QA[x] = ((x / k) == (x % k)) ? 1.0 : 0.0;
        YB_THREADS_END
        *fail_out = (anorm == 0.0) ? 0 : 1;
        return QA;
    }
    const int nleaf = (k + LEAF - 1) / LEAF;
    YB_THREADS_BEGIN
    for (int i = tid; i < k; i += nthr) {
        S.sD[i] = d_io[i] / anorm;
        S.se[i] = (i < k - 1) ? e_in[i] / anorm : 0.0;
    }
    for (long long x = tid; x < (long long)k * k; x += nthr) { QA[x] = 0.0; QB[x] = 0.0; }    // different blocks are decoupled
    YB_THREADS_END
    // ---- tear: every cut c = 32, 64, ... lowers the two diagonal entries next to it by |e[c-1]| ----
    YB_THREADS_BEGIN
    for (int q = 1 + tid; q < nleaf; q += nthr) {
        const int c = q * LEAF;
        const double b = fabs(S.se[c - 1]);
        S.sD[c - 1] -= b;
        S.sD[c] -= b;
    }
    YB_THREADS_END
    // ---- leaves (one warp each) ----
    YB_THREADS_BEGIN
    const int warp = tid >> 5, lane = tid & 31, nwarps = nthr >> 5;
    for (int q = warp; q < nleaf; q += nwarps) {
        const int a = q * LEAF, n = min(LEAF, k - a);
        double dl[LEAF], el[LEAF], z[LEAF];
        for (int i = 0; i < LEAF; ++i) {
            dl[i] = (i < n) ? S.sD[a + i] : 0.0;
            el[i] = (i < n - 1) ? S.se[a + i] : 0.0;
        }
        const int bad = leaf_ql(dl, el, z, n, lane);
        if (bad) S.info[3] = 1;
        if (lane < n) {
            S.Ds[a + lane] = dl[lane];
#pragma unroll
            for (int i = 0; i < LEAF; ++i)
                if (i < n) QA[(long long)(a + i) * k + a + lane] = z[i];
        }
    }
    YB_THREADS_END
    YB_THREADS_BEGIN
    for (int i = tid; i < k; i += nthr) S.sD[i] = S.Ds[i];
    YB_THREADS_END

    double* Qs = QA;      // source of the current level
    double* Qd = QB;
    for (int bs = LEAF; bs < k; bs <<= 1) {
        for (int a = 0; a < k; a += 2 * bs) {
            const int c = a + bs;
            const int b = min(a + 2 * bs, k);
            const int n = b - a;
            if (c >= k) {
                // unpaired block: carried to the next level unchanged
                YB_THREADS_BEGIN
                for (int x = tid; x < n * n; x += nthr) {
                    const long long o = (long long)(a + x / n) * k + a + (x % n);
                    Qd[o] = Qs[o];
                }
                YB_THREADS_END
                continue;
            }
            const int n1 = c - a;
            const double beta = S.se[c - 1];
            const double rho = 2.0 * fabs(beta);
            const double sgn = (beta < 0.0) ? -1.0 : 1.0;
            // ---- z = Q^T u / sqrt(2), maxima for the deflation tolerance ----
            YB_THREADS_BEGIN
            for (int t = tid; t < n; t += nthr) {
                const double zz = (t < n1) ? Qs[(long long)(a + t) * k + (c - 1)] : sgn * Qs[(long long)(a + t) * k + c];
                S.zk[t] = zz * 0.7071067811865476;        // unsorted copies live in zk / dk until the sort
                S.dk[t] = S.sD[a + t];
            }
            YB_THREADS_END
            YB_THREADS_BEGIN
            if (tid < 32) {
                double m = 0.0;
                for (int t = tid; t < n; t += 32) {
                    const double x = fabs(S.dk[t]), y = fabs(S.zk[t]);
                    m = fmax(m, (x == x) ? x : INFINITY);
                    m = fmax(m, (y == y) ? y : INFINITY);
                }
                S.red[tid] = m;
            }
            // rank sort (stable): position of element t among the n poles
            for (int t = tid; t < n; t += nthr) {
                const double v = S.dk[t];
                int r = 0;
                for (int u = 0; u < n; ++u) {
                    const double w = S.dk[u];
                    r += (w < v || (w == v && u < t)) ? 1 : 0;
                }
                S.org[t] = r;                              // scratch: rank
            }
            YB_THREADS_END
            YB_THREADS_BEGIN
            for (int t = tid; t < n; t += nthr) {
                const int r = S.org[t];
                S.perm[r] = t;
                S.Ds[r] = S.dk[t];
                S.zs[r] = S.zk[t];
            }
            if (tid == 0) {
                double m = 0.0;
                for (int i = 0; i < 32; ++i) m = fmax(m, S.red[i]);
                S.red[32] = m;
            }
            YB_THREADS_END
            const double vmax = S.red[32];
            if (!(vmax < INFINITY)) {                       // non-finite data: ranks are meaningless -> flag and stop
                *fail_out = 1;
                YB_THREADS_BEGIN
                for (long long x = tid; x < (long long)k * k; x += nthr) QA[x] = ((x / k) == (x % k)) ? 1.0 : 0.0;
                YB_THREADS_END
                return QA;
            }
            const double tol = 8.0 * EPS * vmax;
            // ---- deflation scan (sequential over the sorted poles) ----
            YB_THREADS_BEGIN
            if (tid == 0) {
                int nk = 0, nd = 0, nr = 0, pj = -1;
                for (int j = 0; j < n; ++j) {
                    if (rho * fabs(S.zs[j]) <= tol) { S.defll[nd++] = j; continue; }
                    if (pj < 0) { pj = j; continue; }
                    const double zp = S.zs[pj], zj = S.zs[j];
                    const double t = S.Ds[j] - S.Ds[pj];
                    const double tau2 = zj * zj + zp * zp;
                    // |t c s| <= tol  with  c = zj / tau, s = -zp / tau
                    if (fabs(t * zj * zp) <= tol * tau2) {
                        const double tau = sqrt(tau2);
                        const double cc = zj / tau, ss = -zp / tau;
                        S.zs[j] = tau;
                        S.zs[pj] = 0.0;
                        S.ra[nr] = pj; S.rb[nr] = j; S.rc[nr] = cc; S.rs[nr] = ss; ++nr;
                        const double dp = S.Ds[pj], dj = S.Ds[j];
                        S.Ds[pj] = dp * cc * cc + dj * ss * ss;
                        S.Ds[j] = dp * ss * ss + dj * cc * cc;
                        S.defll[nd++] = pj;
                        pj = j;
                    } else {
                        S.keepl[nk++] = pj;
                        pj = j;
                    }
                }
                if (pj >= 0) S.keepl[nk++] = pj;
                S.info[0] = nk; S.info[1] = nd; S.info[2] = nr;
            }
            YB_THREADS_END
            const int K = S.info[0], nd = S.info[1], nrot = S.info[2];
            // ---- rotations of close poles on the rows of the source (thread = component), gather the secular problem ----
            YB_THREADS_BEGIN
            if (nrot > 0) {
                for (int t = tid; t < n; t += nthr) {
                    for (int r = 0; r < nrot; ++r) {
                        double* pa = Qs + (long long)(a + S.perm[S.ra[r]]) * k + a + t;
                        double* pb = Qs + (long long)(a + S.perm[S.rb[r]]) * k + a + t;
                        const double qa = *pa, qb = *pb, cc = S.rc[r], ss = S.rs[r];
                        *pa = cc * qa + ss * qb;
                        *pb = cc * qb - ss * qa;
                    }
                }
            }
            for (int i = tid; i < K; i += nthr) {
                const int j = S.keepl[i];
                S.dk[i] = S.Ds[j];
                S.zk[i] = S.zs[j];
                S.krow[i] = S.perm[j];
            }
            YB_THREADS_END
            // ---- secular equation, one thread per root ----
            YB_THREADS_BEGIN
            for (int i = tid; i < K; i += nthr) {
                int o; double m;
                if (secular_root(i, K, S.dk, S.zk, rho, o, m)) S.info[3] = 1;
                S.org[i] = o;
                S.mu[i] = m;
            }
            YB_THREADS_END
            // ---- Gu/Eisenstat: zhat_j^2 = prod_i (lam_i - d_j) / (rho prod_{i != j} (d_i - d_j)) ----
            YB_THREADS_BEGIN
            for (int j = tid; j < K; j += nthr) {
                const double dj = S.dk[j];
                // four interleaved partial products: the quotients are independent, only the running product is a dependent chain, and
                // an FP64 division is ~200 cycles of latency -- one chain of K of them kept the whole CTA at the barrier below for 18 %
                // of the kernel (profiles/r02_ncu_stedc_source_hotspots_v1.txt)
                double pp[4] = {-((dj - S.dk[S.org[j]]) - S.mu[j]) / rho, 1.0, 1.0, 1.0};
                for (int i0 = 0; i0 < K; i0 += 4) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int i = i0 + u;
                        if (i >= K || i == j) continue;
                        const double del = (dj - S.dk[S.org[i]]) - S.mu[i];      // d_j - lam_i
                        pp[u] *= -del / (S.dk[i] - dj);
                    }
                }
                const double p = (pp[0] * pp[1]) * (pp[2] * pp[3]);
                S.zh[j] = copysign(sqrt(fabs(p)), S.zk[j]);
            }
            YB_THREADS_END
            // ---- 1 / norm of every eigenvector of the rank-one problem ----
            YB_THREADS_BEGIN
            for (int i = tid; i < K; i += nthr) {
                const double dorg = S.dk[S.org[i]], m = S.mu[i];
                double s4[4] = {0.0, 0.0, 0.0, 0.0};
                for (int j0 = 0; j0 < K; j0 += 4) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int j = j0 + u;
                        if (j >= K) continue;
                        const double v = S.zh[j] / ((S.dk[j] - dorg) - m);
                        s4[u] += v * v;
                    }
                }
                S.vn[i] = 1.0 / sqrt((s4[0] + s4[1]) + (s4[2] + s4[3]));
            }
            YB_THREADS_END
            // ---- Qd[a + i][a + cc] = sum_j V[i][j] Qs[a + krow[j]][a + cc],  V[i][j] = vn_i zhat_j / (d_j - lam_i) ----
            for (int i0 = 0; i0 < K; i0 += TM) {
                for (int c0 = 0; c0 < n; c0 += TN) {
#ifndef YB_HOST_EMU
                    double acc[32];
#endif
                    YB_THREADS_BEGIN
#pragma unroll
                    for (int r = 0; r < 32; ++r) YB_DC_ACC(r) = 0.0;
                    YB_THREADS_END
                    for (int j0 = 0; j0 < K; j0 += TK) {
                        YB_THREADS_BEGIN
                        for (int x = tid; x < TM * TK; x += nthr) {
                            const int row = x % TM, kk = x / TM;
                            const int i = i0 + row, j = j0 + kk;
                            double v = 0.0;
                            if (i < K && j < K) v = S.vn[i] * S.zh[j] / ((S.dk[j] - S.dk[S.org[i]]) - S.mu[i]);
                            S.As[kk][row] = v;
                        }
                        for (int x = tid; x < TN * TK; x += nthr) {
                            const int cc = x % TN, kk = x / TN;
                            const int j = j0 + kk, col = c0 + cc;
                            double v = 0.0;
                            if (j < K && col < n) v = Qs[(long long)(a + S.krow[j]) * k + a + col];
                            S.Bs[kk][cc] = v;
                        }
                        YB_THREADS_END
                        YB_THREADS_BEGIN
                        const int ty = tid >> 4, tx = tid & 15;          // rows ty*4 .. +3 ; cols tx*2 + 32 u (+0, +1)
#pragma unroll
                        for (int kk = 0; kk < TK; ++kk) {
                            double av[4], bv[8];
#pragma unroll
                            for (int r = 0; r < 4; ++r) av[r] = S.As[kk][ty * 4 + r];
#pragma unroll
                            for (int u = 0; u < 4; ++u) { bv[2 * u] = S.Bs[kk][tx * 2 + 32 * u]; bv[2 * u + 1] = S.Bs[kk][tx * 2 + 32 * u + 1]; }
#pragma unroll
                            for (int r = 0; r < 4; ++r)
#pragma unroll
                                for (int q = 0; q < 8; ++q) YB_DC_ACC(r * 8 + q) = fma(av[r], bv[q], YB_DC_ACC(r * 8 + q));
                        }
                        YB_THREADS_END
                    }
                    YB_THREADS_BEGIN
                    const int ty = tid >> 4, tx = tid & 15;
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const int i = i0 + ty * 4 + r;
                        if (i < K) {
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                const int col = c0 + tx * 2 + 32 * (q >> 1) + (q & 1);
                                if (col < n) Qd[(long long)(a + i) * k + a + col] = YB_DC_ACC(r * 8 + q);
                            }
                        }
                    }
                    YB_THREADS_END
                }
            }
            // ---- deflated eigenpairs are carried over; new eigenvalues ----
            YB_THREADS_BEGIN
            for (int x = tid; x < nd * n; x += nthr) {
                const int t = x / n, col = x % n;
                Qd[(long long)(a + K + t) * k + a + col] = Qs[(long long)(a + S.perm[S.defll[t]]) * k + a + col];
            }
            for (int i = tid; i < K; i += nthr) S.sD[a + i] = S.dk[S.org[i]] + S.mu[i];
            for (int t = tid; t < nd; t += nthr) S.sD[a + K + t] = S.Ds[S.defll[t]];
            YB_THREADS_END
        }
        double* tmp = Qs; Qs = Qd; Qd = tmp;
    }
    YB_THREADS_BEGIN
    for (int i = tid; i < k; i += nthr) d_io[i] = S.sD[i] * anorm;
    YB_THREADS_END
    *fail_out = S.info[3];
    return Qs;
}

}  // namespace stedc
