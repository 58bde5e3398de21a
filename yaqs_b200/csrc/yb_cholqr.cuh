// Truncation-free gauge shifts by Cholesky-QR with a keep-all certificate (round 2, second session).
//
// Where: every SVD gauge shift of the reference whose truncation policy is the hard-coded 1e-12 discarded weight (mps.py:757-759:
// apply_dissipation's two sweeps, dissipation.py:88-183, and normalize("B", "SVD") after a jump, stochastic_process.py:291) or no
// truncation at all (the QR moves, mps.py:736-746).  truncate() (core/linalg/svd_utils.py:60-75) keeps EVERY singular value iff the
// smallest one satisfies sigma_min^2 >= threshold; the factors of a kept-all SVD are then just one orthonormal basis of the column (row)
// space times the complementary factor -- any QR / LQ factorisation is the same MPS up to a unitary on the bond, which no observable,
// no later truncation and no jump probability sees.  In the measured C3 step two thirds of all Gram factorisations were such shifts
// (790 of 1190 per lane-step), each paying a full Hermitian eigen-solve (tridiagonalisation + divide & conquer + back-transformation,
// one SM per matrix, ~2 ms per launch at k = 128) plus a second-level solve for the 1e-12 threshold.
//
// Route (k = the side that becomes the bond, len = the other one, W = k x len work matrix of k_svd_prepare, G = W W^H):
//   k_cholqr_factor    G = L~ D L~^H in shared memory (packed lower triangle, one CTA per trajectory), L^-1 = D^-1/2 L~^-1 in place,
//                      CERTIFICATE  lambda_min(G) >= 1 / ||L^-1||_F^2 ;  a trajectory is accepted iff that bound, less the rounding
//                      allowance of the Gram product and the factorisation, clears 4 x threshold.  Everything else is FLAGGED and
//                      redone by the eigen-solver route.
//   yb_gemm ROTATE     Y1 = L^-1 W            (rows orthonormal up to eps * cond^2 -- the known weakness of Cholesky-QR)
//   yb_gemm GRAM       G2 = Y1 Y1^H = I + E
//   k_cholqr_refine    X = (3 I - G2) / 2 = (I + E)^-1/2 + O(E^2)  (one Newton-Schulz step);  ||E||_F > tol flags the trajectory
//   yb_gemm ROTATE     Y2 = X Y1              (orthonormal to 0.75 ||E||^2 <= 1e-15)
//   yb_gemm GRAM       L2 = W Y2^H            (the exact complementary factor: W = L2 Y2 because Y2 spans the row space of W)
//   k_cholqr_finalize  keep = k;  orient 1 (W = M^H): U = Y2^H, "S Vh" = L2^H;  orient 0 (W = M): "U S" = L2, Vh = Y2.
// All products run on the DMMA kernel with several CTAs per trajectory; only the k^3/3 factorisation is serial per trajectory and it
// works out of shared memory (~50 us instead of ~2 ms).
#pragma once

namespace cholqr {

constexpr int THREADS = 512;
constexpr int KMAX = 128;                                   // packed lower triangle of a 128 x 128 complex matrix: 132 KB of shared memory
constexpr size_t smem_bytes(int kmax) { return ((size_t)kmax * (kmax + 1) / 2 + kmax) * sizeof(cplx) + (size_t)kmax * sizeof(double) + 64 * sizeof(double); }

__device__ __forceinline__ int tri(int i) { return (i * (i + 1)) >> 1; }

// meta: the block of k_svd_prepare (0: k, 1: len).  linv: k x k row-major, zeros above the diagonal.  flag[b] = 1: not accepted.
__global__ void __launch_bounds__(THREADS, 1) k_cholqr_factor(const cplx* __restrict__ G, long long sG, cplx* Linv, long long sL, const int* meta,
                                                              double threshold, int* flag, int* n_flag, int kmax, const int* mask) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_bad;
    const int b = blockIdx.x;
    const int tid = threadIdx.x;
    if (mask && !mask[b]) { if (tid == 0) flag[b] = 0; return; }
    const int k = meta[b * META + 0], len = meta[b * META + 1];
    if (k > len || k > kmax || k < 1) {                      // the bond side must be the short side (otherwise G is singular by shape)
        if (tid == 0) { flag[b] = 1; atomicAdd(n_flag, 1); }
        return;
    }
    cplx* P = reinterpret_cast<cplx*>(smem_raw);             // packed lower triangle: element (i, j <= i) at i (i + 1) / 2 + j
    cplx* trow = P + (size_t)kmax * (kmax + 1) / 2;          // [kmax] scaled row of the inversion
    double* w = reinterpret_cast<double*>(trow + kmax);      // [kmax] 1 / pivot
    double* red = w + kmax;                                  // [64]
    const cplx* Gb = G + (long long)b * sG;
    cplx* Lb = Linv + (long long)b * sL;
    const int lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_bad = 0;
    double tr[1] = {0.0};
    for (int e = tid; e < k * k; e += THREADS) {
        const int i = e / k, j = e - i * k;
        if (j <= i) {
            const cplx g = Gb[e];
            P[tri(i) + j] = g;
            if (i == j) tr[0] += g.x;
        }
    }
    block_sum<1>(tr, red);                                   // (contains the barrier that publishes P)
    const double trace = tr[0];
    // ---- G = L~ D L~^H, left-looking by columns.  S[r][j] = G[r][j] - sum_{q<j} S[r][q] conj(S[j][q]) w[q], pivot_j = S[j][j], w = 1 / pivot.
    // PT threads per row (q = part, part + PT, ...): 4 while more than 64 rows are left, then 8 / 16 / 32 as the triangle narrows, so that
    // the 512 threads stay busy (with a fixed split 43 % of the stall samples were warps waiting at the column / row barriers,
    // profiles/r02_ncu_cholqr_factor_source_hotspots_v1.txt).  One barrier per column: the reciprocal of the newest pivot is recomputed by
    // every thread that needs it and published through w[] for the columns after the next barrier.
    for (int j = 0; j < k; ++j) {
        const int left = k - j;
        const int PT = left > 64 ? 4 : (left > 32 ? 8 : (left > 16 ? 16 : 32));
        const int part = tid & (PT - 1), rslot = tid / PT;
        double wlast = 0.0;
        if (j > 0) {
            const double pv = P[tri(j - 1) + (j - 1)].x;
            wlast = (pv > 0.0 && isfinite(pv)) ? 1.0 / pv : 0.0;
            if (tid == 0) { w[j - 1] = wlast; if (wlast == 0.0) s_bad = 1; }
        }
        const cplx* sj = P + tri(j);
        for (int r0 = j; r0 < k; r0 += THREADS / PT) {     // block-uniform trip count: every lane takes part in the shuffles
            const int r = r0 + rslot;
            const bool live = r < k;
            const cplx* sr = P + tri(live ? r : j);
            double ar = 0.0, ai = 0.0;
            if (live)
                for (int q = part; q < j; q += PT) {
                    const cplx x = sr[q], y = sj[q];
                    const double wq = (q == j - 1) ? wlast : w[q];
                    ar = fma(x.x * y.x + x.y * y.y, wq, ar);     // x * conj(y) * w
                    ai = fma(x.y * y.x - x.x * y.y, wq, ai);
                }
            for (int o = 1; o < PT; o <<= 1) { ar += __shfl_xor_sync(0xffffffffu, ar, o); ai += __shfl_xor_sync(0xffffffffu, ai, o); }
            if (live && part == 0) {
                cplx g = sr[j];
                g.x -= ar; g.y -= ai;
                if (r == j) g.y = 0.0;
                P[tri(r) + j] = g;                           // only this thread group touches (r, j) in this phase; row j's earlier entries are read-only
            }
        }
        __syncthreads();
    }
    {
        const double pv = P[tri(k - 1) + (k - 1)].x;
        const double wl = (pv > 0.0 && isfinite(pv)) ? 1.0 / pv : 0.0;
        if (tid == 0) { w[k - 1] = wl; if (wl == 0.0) s_bad = 1; }
    }
    __syncthreads();
    // ---- Z = L~^-1 in place, row by row: Z[i][j] = - sum_{q=j}^{i-1} (S[i][q] w[q]) Z[q][j], Z[i][i] = 1.  Lane layout inside a warp:
    // JW = 32 / PT columns per warp, column j = JW warp + (lane % JW), part = lane / JW (q = j + part, j + part + PT, ...): lanes of one
    // part read JW consecutive Z[q][.]; PT follows the row length as above (4 for i > 64 ... 32 for i <= 16).
    for (int i = 1; i < k; ++i) {
        for (int q = tid; q < i; q += THREADS) { const cplx s = P[tri(i) + q]; const double wq = w[q]; trow[q] = make_double2(s.x * wq, s.y * wq); }
        __syncthreads();
        const int PT = i > 64 ? 4 : (i > 32 ? 8 : (i > 16 ? 16 : 32));
        const int JW = 32 / PT;
        const int j = JW * warp + (lane & (JW - 1)), pr = lane / JW;
        if (JW * warp < i) {
            double ar = 0.0, ai = 0.0;
            if (j < i) {
                for (int q = j + pr; q < i; q += PT) {
                    const cplx t = trow[q];
                    const cplx z = (q == j) ? make_double2(1.0, 0.0) : P[tri(q) + j];
                    ar += t.x * z.x - t.y * z.y;
                    ai += t.x * z.y + t.y * z.x;
                }
            }
            for (int o = JW; o < 32; o <<= 1) { ar += __shfl_xor_sync(0xffffffffu, ar, o); ai += __shfl_xor_sync(0xffffffffu, ai, o); }
            if (pr == 0 && j < i) P[tri(i) + j] = make_double2(-ar, -ai);      // row i of S is dead: its scaled copy lives in trow
        }
        __syncthreads();
    }
    // ---- L^-1 = D^-1/2 Z (row scaling), its squared Frobenius norm, the certificate, the dense output ---------------------------------
    double fn[1] = {0.0};
    for (int e = tid; e < k * k; e += THREADS) {
        const int i = e / k, j = e - i * k;
        cplx v = make_double2(0.0, 0.0);
        if (j <= i) {
            const double sc = sqrt(w[i]);
            const cplx z = (j == i) ? make_double2(1.0, 0.0) : P[tri(i) + j];
            v = make_double2(z.x * sc, z.y * sc);
            fn[0] += v.x * v.x + v.y * v.y;
        }
        Lb[e] = v;
    }
    block_sum<1>(fn, red);
    if (tid == 0) {
        // lambda_min(L L^H) >= 1 / ||L^-1||_F^2 (= 1 / trace(G^-1));  L L^H differs from the exact W W^H by the rounding of the Gram
        // product and of the factorisation, both <= c k eps trace(G): 64 k eps trace is the allowance.  Whether the condition number
        // is one the single refinement step can repair is not decided here: k_cholqr_refine measures the actual defect.
        const double lb = 1.0 / fn[0] - 64.0 * k * 2.220446049250313e-16 * trace;
        const bool ok = !s_bad && isfinite(fn[0]) && isfinite(trace) && trace > 0.0 && lb >= 4.0 * threshold && lb > 0.0;
        flag[b] = ok ? 0 : 1;
        if (!ok) atomicAdd(n_flag, 1);
    }
}

// X = (3 I - G2) / 2 with G2 = Y1 Y1^H; trajectories whose defect ||G2 - I||_F exceeds tol (or is not finite) are flagged.
__global__ void __launch_bounds__(THREADS) k_cholqr_refine(const cplx* __restrict__ G2, long long sG, cplx* X, long long sX, const int* meta, double tol,
                                                            int* flag, int* n_flag, const int* mask) {
    __shared__ double red[64];
    const int b = blockIdx.x;
    if (mask && !mask[b]) return;
    if (flag[b]) return;
    const int k = meta[b * META + 0];
    const cplx* g = G2 + (long long)b * sG;
    cplx* x = X + (long long)b * sX;
    double e2[1] = {0.0};
    for (int e = threadIdx.x; e < k * k; e += THREADS) {
        const int i = e / k, j = e - i * k;
        cplx v = g[e];
        const double dr = v.x - (i == j ? 1.0 : 0.0);
        e2[0] += dr * dr + v.y * v.y;
        x[e] = make_double2((i == j ? 1.5 : 0.0) - 0.5 * v.x, -0.5 * v.y);
    }
    block_sum<1>(e2, red);
    if (threadIdx.x == 0 && !(e2[0] <= tol * tol)) { flag[b] = 1; atomicAdd(n_flag, 1); }
}

// Factors of the accepted trajectories in the layout k_svd_finalize writes (U: rows x keep, "Vh": keep x cols, both row-major).
__global__ void k_cholqr_finalize(const cplx* __restrict__ Y2, long long sY, const cplx* __restrict__ L2, long long sL, int* meta, int* keep_out,
                                  int keep_stride, cplx* uout, long long s_u, cplx* vout, long long s_v, const int* flag, const int* mask) {
    const int b = blockIdx.y;
    if (mask && !mask[b]) return;
    int* m = meta + b * META;
    if (blockIdx.x == 0 && threadIdx.x == 0) m[7] = 0;       // scale word is per factorisation (k_svd_prepare accumulates it)
    if (flag[b]) return;
    const int k = m[0], len = m[1], orient = m[2];
    if (blockIdx.x == 0 && threadIdx.x == 0) { keep_out[(long long)b * keep_stride] = k; m[5] = 1; }
    const cplx* y = Y2 + (long long)b * sY;
    const cplx* l = L2 + (long long)b * sL;
    cplx* ub = uout + (long long)b * s_u;
    cplx* vb = vout + (long long)b * s_v;
    const long long nbig = (long long)k * len, nsm = (long long)k * k;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nbig + nsm; e += (long long)gridDim.x * blockDim.x) {
        if (e < nbig) {
            if (orient) {                                    // U[j][c] = conj(Y2[c][j]), j < len = rows
                const int j = (int)(e / k), c = (int)(e - (long long)j * k);
                ub[e] = cconj(y[(long long)c * len + j]);
            } else {                                         // Vh = Y2 (k x cols = len)
                vb[e] = y[e];
            }
        } else {
            const long long f = e - nbig;
            const int i = (int)(f / k), j = (int)(f - (long long)i * k);
            if (orient) vb[f] = cconj(l[(long long)j * k + i]);      // "S Vh" = L2^H (k x cols = k)
            else ub[f] = l[f];                                         // "U S" = L2 (rows = k x k)
        }
    }
}

}  // namespace cholqr
