// Batched truncated SVD by one-sided (Hestenes) Jacobi -- replaces scipy.linalg.svd (zgesdd) + linalg.truncate at
// core/linalg/svd.py:51-104, core/linalg/svd_utils.py:22-104, core/methods/decompositions.py:105-185.
//
// Why Jacobi: the reference's gauge shifts truncate at a discarded weight of 1e-12, i.e. they resolve singular values
// down to ~1e-6 relative; one-sided Jacobi delivers high RELATIVE accuracy for every singular value, a Gram-matrix
// eigen-solve does not (DESIGN.md "SVD").  The vectors being orthogonalised are the contiguous rows of W (the smaller
// side of the matrix), so every access is a coalesced 16-byte stream; a CTA owns one block pair of one trajectory,
// keeps the 2*block_rows rows of W and of the accumulated rotation Q in shared memory and sweeps all row pairs with
// one warp per pair.  A full cyclic sweep = (nblocks-1) rounds of a round-robin over block pairs.
#include "yb_common.cuh"
#include <math.h>

// meta[b][8]: 0 k (rows of W), 1 len (row length), 2 orient (0: W = M, 1: W = M^H), 3 rows(M), 4 cols(M), 5 done, 6 sweeps,
//             7 scale: SCALE_BIAS + ilogb(max |W_ij|^2), 0 = unset (written by atomicMax in k_svd_prepare, cleared by k_svd_finalize)
constexpr int META = 8;
constexpr int SVD_KMAX = 1024;
constexpr int SCALE_BIAS = 4096;

// Noise floor of a trajectory's matrix (squared): rows whose squared norm is below tol^2 * 2^ilogb(max |W_ij|^2) <= (tol * sigma_max)^2
// are rounding residue of earlier rotations (rank-deficient input, e.g. a state right after a lowering-operator jump).  They take no
// further part in the sweeps and count as exact zeros in k_svd_finalize.  The rule is ABSOLUTE per matrix: the pairwise rule it
// replaces ("skip a pair whose norms differ by more than 1/eps") was not transitive -- a residue row could be skipped against a
// large row while its equally small neighbour was not, and rotations among the residue rows kept re-mixing the component that
// was never removed; the block schedule then needed > 60 sweeps on 3 of 256 trajectories of the C3 run
// (tests/golden/jacobi/rankdef_c3_*.npy are those matrices).
__device__ __forceinline__ double noise_floor2(const int* m, double tol2) {
    const int sc = m[7];
    return sc > 0 ? ldexp(tol2, sc - SCALE_BIAS) : 0.0;
}
__device__ __forceinline__ double jacobi_tol2(int len) {
    const double tolr = fmax(10.0, sqrt((double)len)) * 2.220446049250313e-16;
    return tolr * tolr;
}

__global__ void k_svd_prepare(int mode, int dphys, const cplx* __restrict__ src, long long s_src, cplx* w, long long s_w, cplx* q, long long s_q,
                              int* meta, int* rot_count, const int* el, int esl, const int* er, int esr, const int* mask, int* n_zero,
                              int* status_zero) {
    const int b = blockIdx.y;
    if (n_zero && b == 0 && blockIdx.x == 0 && threadIdx.x == 0) *n_zero = 0;     // pending / flagged counter of the factorisation that starts here
    if (status_zero && blockIdx.x == 0 && threadIdx.x == 0) status_zero[b] = 0;   // per-trajectory scratch status of the Gram route
    if (mask && !mask[b]) { if (blockIdx.x == 0 && threadIdx.x == 0) meta[b * META + 5] = 1; return; }
    const int l = el[(long long)b * esl], r = er[(long long)b * esr];
    const int force = (mode >> 4) & 3;          // 0: orient by shape, 1: W = M, 2: W = M^H (Gram path picks the isometric side)
    mode &= 3;
    int rows, cols;
    if (mode == 0) { rows = dphys * l; cols = dphys * r; }
    else if (mode == 1) { rows = dphys * l; cols = r; }
    else { rows = l; cols = dphys * r; }
    const int orient = force == 1 ? 0 : (force == 2 ? 1 : (rows <= cols ? 0 : 1));
    const int k = orient ? cols : rows, len = orient ? rows : cols;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        int* m = meta + b * META;
        m[0] = k; m[1] = len; m[2] = orient; m[3] = rows; m[4] = cols; m[5] = 0; m[6] = 0;      // m[7]: see META
        rot_count[b] = 0;
    }
    const cplx* sb = src + (long long)b * s_src;
    cplx* wb = w + (long long)b * s_w;
    cplx* qb = q + (long long)b * s_q;
    const long long nw = (long long)k * len;
    double vmax = 0.0;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nw; e += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(e / len), jj = (int)(e % len);
        const int row = orient ? jj : i, col = orient ? i : jj;           // element M[row][col]
        long long idx;
        if (mode == 0) {            // theta[(s*d+s'), l, r]; row = s*l + li, col = s'*r + ri
            const int s = row / l, li = row % l, sp = col / r, ri = col % r;
            idx = (((long long)(s * dphys + sp)) * l + li) * r + ri;
        } else if (mode == 1) {     // (s l) x r contiguous
            idx = (long long)row * r + col;
        } else {                    // l x (s r): src[(s*l + li)*r + ri]
            const int s = col / r, ri = col % r;
            idx = ((long long)s * l + row) * r + ri;
        }
        cplx v = sb[idx];
        if (orient) v.y = -v.y;
        wb[e] = v;
        vmax = fmax(vmax, cabs2(v));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) vmax = fmax(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    if ((threadIdx.x & 31) == 0 && vmax > 0.0 && isfinite(vmax)) atomicMax(&meta[b * META + 7], SCALE_BIAS + ilogb(vmax));
    const long long nq = (long long)k * k;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nq; e += (long long)gridDim.x * blockDim.x)
        qb[e] = make_double2((e / k) == (e % k) ? 1.0 : 0.0, 0.0);
}

// One pair rotation by a full warp.  wi/wj: rows of W (length len) in shared memory, qi/qj: rows of Q (length k).
__device__ __forceinline__ int jacobi_pair(cplx* wi, cplx* wj, int len, cplx* qi, cplx* qj, int k, double tol2, double floor2, int lane) {
    double a = 0.0, bb = 0.0, cr = 0.0, ci = 0.0;
    for (int e = lane; e < len; e += 32) {
        const cplx x = wi[e], y = wj[e];
        a += x.x * x.x + x.y * x.y;
        bb += y.x * y.x + y.y * y.y;
        cr += x.x * y.x + x.y * y.y;        // conj(x) * y
        ci += x.x * y.y - x.y * y.x;
    }
    a = warp_sum(a); bb = warp_sum(bb); cr = warp_sum(cr); ci = warp_sum(ci);
    const double c2 = cr * cr + ci * ci;
    if (!(c2 > tol2 * a * bb) || c2 == 0.0) return 0;
    // rows below the matrix's noise floor are deflated (noise_floor2): their singular values stay <= tol * sigma_max, i.e. within
    // the absolute accuracy zgesdd gives the reference
    if (a < floor2 || bb < floor2) return 0;
    const double absc = sqrt(c2);
    const double zeta = (bb - a) / (2.0 * absc);
    const double t = fabs(zeta) > 1e100 ? 0.5 / zeta : copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
    const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
    // wj~ = conj(e) wj with e = c/|c|;  wi' = cs wi - sn wj~ ;  wj' = sn wi + cs wj~
    const double er = cr / absc, ei = -ci / absc;     // conj(e)
    for (int e = lane; e < len; e += 32) {
        const cplx x = wi[e], y = wj[e];
        const double yr = er * y.x - ei * y.y, yi = er * y.y + ei * y.x;
        wi[e] = make_double2(cs * x.x - sn * yr, cs * x.y - sn * yi);
        wj[e] = make_double2(sn * x.x + cs * yr, sn * x.y + cs * yi);
    }
    for (int e = lane; e < k; e += 32) {
        const cplx x = qi[e], y = qj[e];
        const double yr = er * y.x - ei * y.y, yi = er * y.y + ei * y.x;
        qi[e] = make_double2(cs * x.x - sn * yr, cs * x.y - sn * yi);
        qj[e] = make_double2(sn * x.x + cs * yr, sn * x.y + cs * yi);
    }
    return 1;
}

// grid = (nblocks/2, B); block = 512 threads; dynamic smem = 2*block_rows*(max_len + max_k) complex.
__global__ void __launch_bounds__(512) k_svd_round(int block_rows, int nblocks, int round, int full_panel, int inner_sweeps, cplx* w, long long s_w,
                                                    cplx* q, long long s_q, int* meta, int* rot_count, int max_len, int max_k, const int* mask,
                                                    int* n_zero) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_rot;
    const int b = blockIdx.y;
    if (n_zero && b == 0 && blockIdx.x == 0 && threadIdx.x == 0) *n_zero = 0;     // counter k_svd_check adds to after the rounds of this sweep
    if (mask && !mask[b]) return;
    const int* m = meta + b * META;
    if (m[5]) return;
    const int k = m[0], len = m[1];
    // round-robin (circle method) over nblocks players
    const int p = blockIdx.x, N1 = nblocks - 1;
    int bi, bj;
    if (p == 0) { bi = N1; bj = round % N1; }
    else { bi = (round + p) % N1; bj = ((round - p) % N1 + N1) % N1; }
    if (bi > bj) { const int t = bi; bi = bj; bj = t; }
    const int i0 = bi * block_rows, j0 = bj * block_rows;
    const int ni = max(0, min(block_rows, k - i0)), nj = max(0, min(block_rows, k - j0));
    if (ni == 0 || (nj == 0 && !full_panel)) return;
    const int nrows = ni + nj;
    cplx* Wp = reinterpret_cast<cplx*>(smem_raw);
    cplx* Qp = Wp + (size_t)2 * block_rows * max_len;
    cplx* wb = w + (long long)b * s_w;
    cplx* qb = q + (long long)b * s_q;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    for (int t = 0; t < nrows; ++t) {
        const int gr = t < ni ? i0 + t : j0 + (t - ni);
        for (int e = tid; e < len; e += blockDim.x) Wp[(size_t)t * max_len + e] = wb[(long long)gr * len + e];
        for (int e = tid; e < k; e += blockDim.x) Qp[(size_t)t * max_k + e] = qb[(long long)gr * k + e];
    }
    if (tid == 0) s_rot = 0;
    __syncthreads();
    const double tol2 = jacobi_tol2(len);
    const double floor2 = noise_floor2(m, tol2);
    int report_rot = 0;
    for (int sweep = 0; sweep < inner_sweeps; ++sweep) {
        int my_rot = 0;
        if (full_panel) {
            // all pairs among nrows rows: circle method over ne = even(nrows) players
            const int ne = (nrows + 1) & ~1, n1 = ne - 1;
            for (int rr = 0; rr < n1; ++rr) {
                for (int pp = warp; pp < ne / 2; pp += nwarps) {
                    int x, y;
                    if (pp == 0) { x = n1; y = rr % n1; } else { x = (rr + pp) % n1; y = ((rr - pp) % n1 + n1) % n1; }
                    if (x < nrows && y < nrows)
                        my_rot += jacobi_pair(Wp + (size_t)min(x, y) * max_len, Wp + (size_t)max(x, y) * max_len, len,
                                              Qp + (size_t)min(x, y) * max_k, Qp + (size_t)max(x, y) * max_k, k, tol2, floor2, lane);
                }
                __syncthreads();
            }
        } else {
            const int mm = max(ni, nj);
            for (int qq = 0; qq < mm; ++qq) {
                for (int t = warp; t < mm; t += nwarps) {
                    const int x = t, y = (t + qq) % mm;
                    if (x < ni && y < nj)
                        my_rot += jacobi_pair(Wp + (size_t)x * max_len, Wp + (size_t)(ni + y) * max_len, len, Qp + (size_t)x * max_k,
                                              Qp + (size_t)(ni + y) * max_k, k, tol2, floor2, lane);
                }
                __syncthreads();
            }
        }
        if (lane == 0 && my_rot) atomicAdd(&s_rot, my_rot);
        __syncthreads();
        report_rot = s_rot;          // rotations of the most recent sweep (0 => this panel is converged)
        __syncthreads();
        if (tid == 0) s_rot = 0;
        __syncthreads();
        if (report_rot == 0) break;
    }
    for (int t = 0; t < nrows; ++t) {
        const int gr = t < ni ? i0 + t : j0 + (t - ni);
        for (int e = tid; e < len; e += blockDim.x) wb[(long long)gr * len + e] = Wp[(size_t)t * max_len + e];
        for (int e = tid; e < k; e += blockDim.x) qb[(long long)gr * k + e] = Qp[(size_t)t * max_k + e];
    }
    if (tid == 0 && report_rot) atomicAdd(&rot_count[b], report_rot);
}

__global__ void k_svd_check(int nbatch, int* meta, int* rot_count, int* n_pending, const int* mask) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nbatch) return;
    if (mask && !mask[b]) return;
    int* m = meta + b * META;
    if (m[5]) return;
    m[6] += 1;
    if (rot_count[b] == 0) m[5] = 1; else atomicAdd(n_pending, 1);
    rot_count[b] = 0;
}

#include "yb_truncate.cuh"


// One CTA per trajectory.
__global__ void __launch_bounds__(512) k_svd_finalize(int trunc_mode, double threshold, int max_bond, int min_keep, int dist, const cplx* __restrict__ w,
                                                       long long s_w, const cplx* __restrict__ q, long long s_q, int* meta, double* sval,
                                                       long long s_s, int* keep_out, int keep_stride, cplx* uout, long long s_u, cplx* vout,
                                                       long long s_v, int* status, const int* mask, const double* __restrict__ lam = nullptr,
                                                       long long s_lam = 0, int top = 0) {
    __shared__ double sig[SVD_KMAX];
    __shared__ double ssort[SVD_KMAX];
    __shared__ int perm[SVD_KMAX];
    __shared__ int s_keep, s_ndef;
    __shared__ double s_tiny;
    __shared__ double red[32];
    const int b = blockIdx.x;
    if (mask && !mask[b]) return;
    const int* m = meta + b * META;
    const int k = m[0], len = m[1], orient = m[2], rows = m[3], cols = m[4];
    const cplx* wb = w + (long long)b * s_w;
    const cplx* qb = q + (long long)b * s_q;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    // "top" mode of the Gram route (yb_heev_rows_top): w / q hold only the rows of the `top` largest eigenvalues (all a split capped at
    // max_bond <= top can keep); the singular values behind them -- needed for the discarded weight only -- are the square roots of the
    // remaining Gram eigenvalues lam[ntop ..) (descending), accurate to the absolute level yb_svd_ambiguous already allows for
    const int ntop = (lam && top > 0) ? min(k, top) : k;
    for (int i = warp; i < ntop; i += nwarps) {
        double s = 0.0;
        for (int e = lane; e < len; e += 32) s += cabs2(wb[(long long)i * len + e]);
        s = warp_sum(s);
        if (lane == 0) sig[i] = sqrt(s);
    }
    for (int i = ntop + tid; i < k; i += blockDim.x) sig[i] = sqrt(fmax(lam[(long long)b * s_lam + i], 0.0));
    __syncthreads();
    for (int i = tid; i < k; i += blockDim.x) {
        const double si = sig[i];
        int rank = i;
        if (i < ntop) {
            rank = 0;
            for (int j = 0; j < ntop; ++j) { const double sj = sig[j]; rank += (sj > si) || (sj == si && j < i); }
        }
        perm[rank] = i; ssort[rank] = si;
    }
    __syncthreads();
    if (tid == 0) {
        s_keep = truncate_keep(ssort, min(k, min(rows, cols)), trunc_mode, threshold, max_bond, min_keep);   // k > min(rows, cols) only on the Gram path
        bool bad = false;
        for (int i = 0; i < k; ++i) bad |= !isfinite(ssort[i]);
        if (bad || !m[5]) atomicOr(&status[b], bad ? 2 : 1);       // 2: non-finite, 1: Jacobi not converged
        int nd = 0;
        const double tiny = fmax(ssort[0] * 4.5e-16, sqrt(noise_floor2(m, jacobi_tol2(len))));
        s_tiny = tiny;
        for (int i = 0; i < s_keep; ++i) nd += !(ssort[i] > tiny);
        s_ndef = nd;
        keep_out[(long long)b * keep_stride] = s_keep;
    }
    __syncthreads();
    if (tid == 0) meta[b * META + 7] = 0;          // scale word is per factorisation (k_svd_prepare accumulates it with atomicMax)
    const int keep = s_keep;
    for (int i = tid; i < k; i += blockDim.x) sval[(long long)b * s_s + i] = ssort[i];
    cplx* ub = uout + (long long)b * s_u;    // rows x keep
    cplx* vb = vout + (long long)b * s_v;    // keep x cols
    const double tiny = s_tiny;                    // deflated rows (see noise_floor2) count as exact zeros
    // Q-side factor (exactly unitary rows of Q) and W-side factor (rows of W / sigma)
    if (orient == 0) {
        // U[:, c] = conj(Q[perm c, :]) * f ;  Vh[c, :] = W[perm c, :] * (g / sigma)
        for (long long e = tid; e < (long long)rows * keep; e += blockDim.x) {
            const int j = (int)(e / keep), c = (int)(e % keep);
            const double s = ssort[c];
            const double f = dist == 0 ? s : (dist == 1 ? 1.0 : sqrt(s));
            const cplx v = qb[(long long)perm[c] * k + j];
            ub[e] = make_double2(v.x * f, -v.y * f);
        }
        for (long long e = tid; e < (long long)keep * cols; e += blockDim.x) {
            const int c = (int)(e / cols), j = (int)(e % cols);
            const double s = ssort[c];
            double sc;
            if (dist == 1) sc = 1.0; else if (!(s > tiny)) sc = 0.0; else sc = (dist == 0 ? 1.0 / s : 1.0 / sqrt(s));
            const cplx v = wb[(long long)perm[c] * len + j];
            vb[e] = make_double2(v.x * sc, v.y * sc);
        }
    } else {
        // U[:, c] = conj(W[perm c, :]) * (f / sigma) ;  Vh[c, :] = Q[perm c, :] * g
        for (long long e = tid; e < (long long)rows * keep; e += blockDim.x) {
            const int j = (int)(e / keep), c = (int)(e % keep);
            const double s = ssort[c];
            double sc;
            if (dist == 0) sc = 1.0; else if (!(s > tiny)) sc = 0.0; else sc = (dist == 1 ? 1.0 / s : 1.0 / sqrt(s));
            const cplx v = wb[(long long)perm[c] * len + j];
            ub[e] = make_double2(v.x * sc, -v.y * sc);
        }
        for (long long e = tid; e < (long long)keep * cols; e += blockDim.x) {
            const int c = (int)(e / cols), j = (int)(e % cols);
            const double s = ssort[c];
            const double g = dist == 0 ? 1.0 : (dist == 1 ? s : sqrt(s));
            const cplx v = qb[(long long)perm[c] * k + j];
            vb[e] = make_double2(v.x * g, v.y * g);
        }
    }
    if (s_ndef == 0) return;
    // Rank-deficient kept vectors (sigma == 0, only reachable through min_keep): LAPACK returns an arbitrary orthonormal
    // completion; we build one by Gram-Schmidt of unit vectors against the vectors already present.  Only needed where
    // the W-side factor must stay an isometry, i.e. when it does not carry S.
    const bool wside_is_v = (orient == 0);
    if ((wside_is_v && dist == 1) || (!wside_is_v && dist == 0)) return;    // W-side carries S: zero vector is exact
    __syncthreads();
    cplx* base = wside_is_v ? vb : ub;
    const long long vstride = wside_is_v ? cols : 1;      // distance between vectors
    const long long estride = wside_is_v ? 1 : keep;      // distance between elements of one vector
    const int vlen = wside_is_v ? cols : rows;
    int next_unit = 0;
    for (int c = 0; c < keep; ++c) {
        if (ssort[c] > tiny) continue;
        cplx* vc = base + (long long)c * vstride;
        for (int attempt = 0; attempt < vlen; ++attempt, ++next_unit) {
            const int u = next_unit % vlen;
            __syncthreads();
            for (int e = tid; e < vlen; e += blockDim.x) vc[(long long)e * estride] = make_double2(e == u ? 1.0 : 0.0, 0.0);
            __syncthreads();
            for (int pass = 0; pass < 2; ++pass)
                for (int o = 0; o < keep; ++o) {
                    if (o == c || (o > c && !(ssort[o] > tiny))) continue;
                    const cplx* vo = base + (long long)o * vstride;
                    double pr[2] = {0.0, 0.0};
                    for (int e = tid; e < vlen; e += blockDim.x) {
                        const cplx x = vo[(long long)e * estride], y = vc[(long long)e * estride];
                        pr[0] += x.x * y.x + x.y * y.y; pr[1] += x.x * y.y - x.y * y.x;
                    }
                    double p0[1] = {pr[0]}, p1[1] = {pr[1]};
                    block_sum<1>(p0, red); block_sum<1>(p1, red);
                    for (int e = tid; e < vlen; e += blockDim.x) {
                        const cplx x = vo[(long long)e * estride];
                        cplx y = vc[(long long)e * estride];
                        y.x -= p0[0] * x.x - p1[0] * x.y; y.y -= p0[0] * x.y + p1[0] * x.x;
                        vc[(long long)e * estride] = y;
                    }
                    __syncthreads();
                }
            double nn[1] = {0.0};
            for (int e = tid; e < vlen; e += blockDim.x) nn[0] += cabs2(vc[(long long)e * estride]);
            block_sum<1>(nn, red);
            if (nn[0] > 0.25) {
                const double inv = 1.0 / sqrt(nn[0]);
                for (int e = tid; e < vlen; e += blockDim.x) { cplx y = vc[(long long)e * estride]; y.x *= inv; y.y *= inv; vc[(long long)e * estride] = y; }
                ++next_unit;
                break;
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Second level of the Gram route.  The eigenvalues of a Gram matrix carry an ABSOLUTE error ~ k eps s_max^2, so after the
// first rotation the rows of Wf whose squared norm is below tau_rel * s_max^2 are only orthogonal to that absolute level.
// They are gathered into a compact block, re-diagonalised through THEIR OWN Gram matrix (whose scale is tau_rel times
// smaller, hence tau_rel times more accurate) and scattered back -- the singular values around the 1e-12 discarded-weight
// threshold of the gauge shifts (mps.py:757-759) come out with ~1e-8 relative accuracy instead of ~1e-3.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) k_svd_select(const cplx* __restrict__ w, long long s_w, const int* meta, double tau_rel, int* sel_n,
                                                     int* sel_idx, int max_k, const int* mask) {
    __shared__ double nu[SVD_KMAX];
    const int b = blockIdx.x;
    if (mask && !mask[b]) { if (threadIdx.x == 0) sel_n[b] = 0; return; }
    const int* m = meta + b * META;
    const int k = m[0], len = m[1];
    const cplx* wb = w + (long long)b * s_w;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    for (int i = warp; i < k; i += nwarps) {
        double s = 0.0;
        for (int e = lane; e < len; e += 32) s += cabs2(wb[(long long)i * len + e]);
        s = warp_sum(s);
        if (lane == 0) nu[i] = s;
    }
    __syncthreads();
    if (tid == 0) {
        double mx = 0.0;
        for (int i = 0; i < k; ++i) mx = fmax(mx, nu[i]);
        int n = 0;
        int* idx = sel_idx + (long long)b * max_k;
        for (int i = 0; i < k; ++i) if (nu[i] < tau_rel * mx) idx[n++] = i;
        sel_n[b] = n >= 2 ? n : 0;
    }
}

// dir = 0: ws[r] = w[idx[r]], qs[r] = q[idx[r]] ;  dir = 1: the reverse.
__global__ void k_svd_gather(int dir, cplx* w, long long s_w, cplx* q, long long s_q, cplx* ws, long long s_ws, cplx* qs, long long s_qs,
                             const int* meta, const int* sel_n, const int* sel_idx, int max_k, const int* mask) {
    const int b = blockIdx.y;
    if (mask && !mask[b]) return;
    const int n = sel_n[b];
    if (n <= 0) return;
    const int* m = meta + b * META;
    const int k = m[0], len = m[1];
    const int* idx = sel_idx + (long long)b * max_k;
    cplx* wb = w + (long long)b * s_w; cplx* qb = q + (long long)b * s_q;
    cplx* wsb = ws + (long long)b * s_ws; cplx* qsb = qs + (long long)b * s_qs;
    const long long nw = (long long)n * len, nq = (long long)n * k;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nw + nq; e += (long long)gridDim.x * blockDim.x) {
        if (e < nw) {
            const int r = (int)(e / len), c = (int)(e % len);
            if (dir == 0) wsb[e] = wb[(long long)idx[r] * len + c]; else wb[(long long)idx[r] * len + c] = wsb[e];
        } else {
            const long long f = e - nw;
            const int r = (int)(f / k), c = (int)(f % k);
            if (dir == 0) qsb[f] = qb[(long long)idx[r] * k + c]; else qb[(long long)idx[r] * k + c] = qsb[f];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Gram route, last step: is a trajectory's keep decision within the error bound of its threshold?  Gram eigenvalues carry an
// absolute error ~ noise_rel * s_max^2 (tau_rel times less for the rows the second level re-diagonalised).  A trajectory is
// AMBIGUOUS -- and is redone by the one-sided Jacobi -- when a cumulative discarded weight (mode 0) or a relative singular value
// (mode 1) lies within that bound of the threshold, when the eigen-solver did not converge, or when the spectrum is not finite.
// One thread per trajectory (the suffix sums run from the smallest value upward, like truncate()); n_amb counts the flagged ones.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void k_svd_ambiguous(int nbatch, const int* meta, const double* __restrict__ sval, long long s_s, int trunc_mode, double threshold,
                                int two_level, double tau_rel, double noise_rel, const int* status, int* amb, int* n_amb, const int* mask) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nbatch) return;
    if (mask && !mask[b]) { amb[b] = 0; return; }
    const int* m = meta + b * META;
    const int nsv = min(m[3], m[4]);
    const double* s = sval + (long long)b * s_s;
    const double smax2 = nsv > 0 ? s[0] * s[0] : 0.0;
    const double err_hi = noise_rel * smax2, err_lo = two_level ? err_hi * tau_rel : err_hi;
    bool flag = (m[5] == 0) || (status && status[b] != 0) || !isfinite(smax2);
    if (trunc_mode == 0) {
        if (threshold > 0.0) {
            double cum = 0.0;
            for (int i = nsv - 1; i >= 0; --i) {
                const double s2 = s[i] * s[i];
                cum += s2;
                const double err = (s2 >= tau_rel * smax2) ? err_hi : err_lo;
                flag |= fabs(cum - threshold) <= err + 1e-9 * threshold;
            }
        }
    } else {
        const double cut = threshold * threshold * smax2;
        for (int i = 0; i < nsv; ++i) {
            const double s2 = s[i] * s[i];
            const double err = (s2 >= tau_rel * smax2) ? err_hi : err_lo;
            flag |= fabs(s2 - cut) <= err;
        }
    }
    amb[b] = flag ? 1 : 0;
    if (flag) atomicAdd(n_amb, 1);
}

#include "yb_cholqr.cuh"
