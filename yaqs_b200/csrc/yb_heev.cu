// Batched Hermitian eigen-solver (one CTA per trajectory) used by the Gram-matrix SVD path:
//   G = W W^H  (yb_gemm GRAM)  ->  G = Qh T Qh^H (k_hetrd, k_ungtr)  ->  T = Z diag(lam) Z^T (k_steqr)
// It replaces the bulk of scipy.linalg.svd (zgesdd) at core/linalg/svd.py:51-104 for the large (2 chi x 2 chi) matrices of
// split_two_site (core/methods/decompositions.py:105-185); the one-sided Jacobi of yb_svd.cu stays the accurate fallback.
//
// All three kernels stream the k x k matrix from L2/HBM with the same access pattern: a warp owns a block of 32
// COLUMNS (lane = column, one coalesced 512-byte row segment per load) and walks down the rows, so the reductions the
// algorithms need (A v by Hermitian symmetry, v^H Q) are per-lane column sums with no shuffles; several row groups per
// column block are combined through shared memory.  Householder update and the next step's matrix-vector product are
// fused into ONE read + ONE write of the trailing block per step (k^3/3 elements in total).
#include "yb_common.cuh"
#include <math.h>
#include "yb_stedc.cuh"
#include "yb_hetrd.cuh"

namespace heev {

constexpr int THREADS = 512;            // 16 warps
constexpr int NWARPS = THREADS / 32;
constexpr int KMAX = 512;               // largest matrix (2 * chi_max for chi_max = 256)
constexpr int RSMAX = 4;                // row groups per column block
constexpr size_t SMEM_BYTES = (size_t)(3 + RSMAX) * KMAX * sizeof(cplx);
// Row ring of the streaming passes: every thread keeps RING rows of its column in flight with cp.async (LDGSTS.128 into a
// warp-private slice of shared memory), so a Householder step is no longer (rows / 4) dependent global-load round trips --
// ncu (profiles/r01_ncu_heev_full_stedc.json) showed 56-62 % of the stall samples of k_hetrd / k_ungtr on those loads with
// DRAM at 52 % of its rate.  A thread only ever reads back its own copies, so cp.async.wait_group is the only synchronisation.
// Depth is a template parameter (0 = direct register loads); yb_heev picks it by matrix size from measurements.
constexpr size_t ring_bytes(int ring) { return (size_t)NWARPS * ring * 32 * sizeof(cplx); }

__device__ __forceinline__ void ring_load(cplx* smem_dst, const cplx* gmem_src) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" :: "r"(dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void ring_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void ring_wait() { asm volatile("cp.async.wait_group %0;\n" :: "n"(N) : "memory"); }

struct Split { int ncb, rsf, cb, rs; bool on; };

__device__ __forceinline__ Split make_split(int ncols) {
    Split s;
    s.ncb = (ncols + 31) >> 5;
    s.rsf = s.ncb > 0 ? max(1, min(RSMAX, NWARPS / s.ncb)) : 1;
    const int wid = threadIdx.x >> 5;
    s.cb = s.ncb > 0 ? wid % s.ncb : 0;
    s.rs = s.ncb > 0 ? wid / s.ncb : 0;
    s.on = s.ncb > 0 && s.rs < s.rsf;
    return s;
}

// zlarfg on x[t] = conj(row[t]), t in [0, ns): v[0] = 1, v[t] = x[t] / (alpha - beta); all threads return (tau, beta).
__device__ __forceinline__ void householder_from_row(const cplx* row, int ns, cplx* v, cplx& tau, double& beta, double* red) {
    double s[1] = {0.0};
    for (int t = 1 + threadIdx.x; t < ns; t += blockDim.x) s[0] += cabs2(row[t]);
    block_sum<1>(s, red);
    const cplx alpha = cconj(row[0]);
    const double xn2 = s[0];
    if (xn2 == 0.0 && alpha.y == 0.0) {
        tau = make_double2(0.0, 0.0);
        beta = alpha.x;
        for (int t = threadIdx.x; t < ns; t += blockDim.x) v[t] = make_double2(t == 0 ? 1.0 : 0.0, 0.0);
    } else {
        beta = -copysign(sqrt(alpha.x * alpha.x + alpha.y * alpha.y + xn2), alpha.x);
        tau = make_double2((beta - alpha.x) / beta, -alpha.y / beta);
        const double dr = alpha.x - beta, di = alpha.y, dn = 1.0 / (dr * dr + di * di);
        const cplx sc = make_double2(dr * dn, -di * dn);           // 1 / (alpha - beta)
        for (int t = threadIdx.x; t < ns; t += blockDim.x) v[t] = (t == 0) ? make_double2(1.0, 0.0) : cmul(cconj(row[t]), sc);
    }
    __syncthreads();
}

// Reduction to real tridiagonal form, lower variant of zhetd2 on full Hermitian storage.  On exit row j of G holds the
// Householder vector v_j in columns j+1.. (v_j[0] = 1), d/e the tridiagonal, tau the reflector scalars.
template <int RING>
__global__ void __launch_bounds__(THREADS, 1) k_hetrd(cplx* G, long long sG, const int* ek, int esk, double* dout, double* eout, cplx* tauout,
                                                       long long s_vec, const int* mask) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double red[64];
    const int b = blockIdx.x;
    if (mask && !mask[b]) return;
    const int n = ek[(long long)b * esk];
    if (n <= 0) return;
    cplx* v = reinterpret_cast<cplx*>(smem_raw);
    cplx* w = v + KMAX;
    cplx* vn = w + KMAX;
    cplx* part = vn + KMAX;                 // [RSMAX][KMAX]
    cplx* ring = part + RSMAX * KMAX;       // [NWARPS][RING][32]   (RING > 0 only)
    cplx* Gb = G + (long long)b * sG;
    double* d = dout + (long long)b * s_vec;
    double* e = eout + (long long)b * s_vec;
    cplx* tauv = tauout + (long long)b * s_vec;
    const int tid = threadIdx.x, lane = tid & 31;
    if (n == 1) { if (tid == 0) { d[0] = Gb[0].x; e[0] = 0.0; } return; }

    cplx tau; double beta;
    householder_from_row(Gb + 1, n - 1, v, tau, beta, red);
    // p_raw = G[S,S] v as column sums: p[c] = sum_i conj(G[i][c]) v[i]
    {
        const Split sp = make_split(n - 1);
        if (sp.on) {
            const int c = 1 + sp.cb * 32 + lane;
            cplx acc = make_double2(0.0, 0.0);
            if (c < n)
                for (int i = 1 + sp.rs; i < n; i += sp.rsf) acc = cfma(cconj(Gb[(long long)i * n + c]), v[i - 1], acc);
            if (c < n) part[sp.rs * KMAX + (c - 1)] = acc;
        }
        __syncthreads();
        for (int t = tid; t < n - 1; t += blockDim.x) {
            cplx acc = part[t];
            for (int r = 1; r < sp.rsf; ++r) acc = cadd(acc, part[r * KMAX + t]);
            w[t] = acc;                      // w holds p_raw for now
        }
        __syncthreads();
    }
    for (int j = 0; j < n - 1; ++j) {
        const int ns = n - 1 - j;            // S = {j+1 .. n-1}
        // p = tau * p_raw ; alpha = -1/2 tau (p^H v) ; w = p + alpha v
        double dot[2] = {0.0, 0.0};
        for (int t = tid; t < ns; t += blockDim.x) {
            const cplx p = cmul(tau, w[t]);
            w[t] = p;
            const cplx vv = v[t];
            dot[0] += p.x * vv.x + p.y * vv.y;       // conj(p) * v
            dot[1] += p.x * vv.y - p.y * vv.x;
        }
        block_sum<2>(dot, red);
        const cplx al = cmul(make_double2(-0.5 * tau.x, -0.5 * tau.y), make_double2(dot[0], dot[1]));
        for (int t = tid; t < ns; t += blockDim.x) w[t] = cfma(al, v[t], w[t]);
        if (tid == 0) { d[j] = Gb[(long long)j * n + j].x; e[j] = beta; tauv[j] = tau; }
        __syncthreads();
        // store v_j into row j, update row j+1 (local index 0) of the trailing block
        {
            cplx* rowj = Gb + (long long)j * n + (j + 1);
            cplx* row1 = Gb + (long long)(j + 1) * n + (j + 1);
            const cplx v0 = v[0], w0 = w[0];
            for (int t = tid; t < ns; t += blockDim.x) {
                rowj[t] = v[t];
                cplx g = row1[t];
                const cplx wc = w[t], vc = v[t];
                // g -= v0 * conj(wc) + w0 * conj(vc)
                g.x -= v0.x * wc.x + v0.y * wc.y + w0.x * vc.x + w0.y * vc.y;
                g.y -= v0.y * wc.x - v0.x * wc.y + w0.y * vc.x - w0.x * vc.y;
                row1[t] = g;
            }
        }
        __syncthreads();
        if (ns == 1) {
            if (tid == 0) { d[n - 1] = Gb[(long long)(n - 1) * n + (n - 1)].x; e[n - 1] = 0.0; }
            break;
        }
        cplx taun; double betan;
        householder_from_row(Gb + (long long)(j + 1) * n + (j + 2), ns - 1, vn, taun, betan, red);
        // fused pass over rows/cols >= j+2: G -= v w^H + w v^H ;  p_next[c] = sum_i conj(G'[i][c]) vn[i]
        {
            const int c0 = j + 2, ncols = n - c0;
            const Split sp = make_split(ncols);
            if (sp.on) {
                const int c = c0 + sp.cb * 32 + lane;
                cplx acc = make_double2(0.0, 0.0);
                if constexpr (RING > 0) {
                  if (c < n) {
                    const cplx wc = w[c - (j + 1)], vc = v[c - (j + 1)];
                    cplx* rg = ring + ((tid >> 5) * RING) * 32 + lane;          // slot s of this thread: rg[s * 32]
                    const int i0 = c0 + sp.rs;
                    const int nr = (n - i0 + sp.rsf - 1) / sp.rsf;                // rows of this thread
                    cplx* src = Gb + (long long)i0 * n + c;
                    const long long step = (long long)sp.rsf * n;
#pragma unroll
                    for (int q = 0; q < RING; ++q) {
                        if (q < nr) ring_load(rg + q * 32, src + q * step);
                        ring_commit();
                    }
                    int slot = 0;
                    for (int r = 0; r < nr; ++r) {
                        ring_wait<RING - 1>();
                        cplx g = rg[slot * 32];
                        const int ii = i0 + r * sp.rsf;
                        const cplx vi = v[ii - (j + 1)], wi = w[ii - (j + 1)];
                        g.x -= vi.x * wc.x + vi.y * wc.y + wi.x * vc.x + wi.y * vc.y;
                        g.y -= vi.y * wc.x - vi.x * wc.y + wi.y * vc.x - wi.x * vc.y;
                        src[r * step] = g;
                        acc = cfma(cconj(g), vn[ii - c0], acc);
                        if (r + RING < nr) ring_load(rg + slot * 32, src + (r + RING) * step);      // after g has been consumed
                        ring_commit();
                        slot = (slot + 1 == RING) ? 0 : slot + 1;
                    }
                    ring_wait<0>();
                    part[sp.rs * KMAX + (c - c0)] = acc;
                  }
                } else if (c < n) {
                    const cplx wc = w[c - (j + 1)], vc = v[c - (j + 1)];
                    int i = c0 + sp.rs;
                    for (; i + 3 * sp.rsf < n; i += 4 * sp.rsf) {
                        cplx g[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) g[u] = Gb[(long long)(i + u * sp.rsf) * n + c];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int ii = i + u * sp.rsf;
                            const cplx vi = v[ii - (j + 1)], wi = w[ii - (j + 1)];
                            g[u].x -= vi.x * wc.x + vi.y * wc.y + wi.x * vc.x + wi.y * vc.y;
                            g[u].y -= vi.y * wc.x - vi.x * wc.y + wi.y * vc.x - wi.x * vc.y;
                            Gb[(long long)ii * n + c] = g[u];
                            acc = cfma(cconj(g[u]), vn[ii - c0], acc);
                        }
                    }
                    for (; i < n; i += sp.rsf) {
                        cplx g = Gb[(long long)i * n + c];
                        const cplx vi = v[i - (j + 1)], wi = w[i - (j + 1)];
                        g.x -= vi.x * wc.x + vi.y * wc.y + wi.x * vc.x + wi.y * vc.y;
                        g.y -= vi.y * wc.x - vi.x * wc.y + wi.y * vc.x - wi.x * vc.y;
                        Gb[(long long)i * n + c] = g;
                        acc = cfma(cconj(g), vn[i - c0], acc);
                    }
                    part[sp.rs * KMAX + (c - c0)] = acc;
                }
            }
            __syncthreads();
            // next step's operands: v <- vn, w <- p_raw_next
            for (int t = tid; t < ncols; t += blockDim.x) {
                cplx acc = part[t];
                for (int r = 1; r < sp.rsf; ++r) acc = cadd(acc, part[r * KMAX + t]);
                w[t] = acc;
                v[t] = vn[t];
            }
            tau = taun; beta = betan;
            __syncthreads();
        }
    }
}

// Q = H_0 H_1 ... H_{n-2} by backward accumulation (zungtr, lower): one fused pass per reflector over the block it touches.
template <int RING>
__global__ void __launch_bounds__(THREADS, RING > 6 ? 1 : 2) k_ungtr(const cplx* __restrict__ G, long long sG, const cplx* __restrict__ tauin, long long s_vec,
                                                       cplx* Q, long long sQ, const int* ek, int esk, const int* mask) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int b = blockIdx.x;
    if (mask && !mask[b]) return;
    const int n = ek[(long long)b * esk];
    if (n <= 0) return;
    cplx* v = reinterpret_cast<cplx*>(smem_raw);    // v_j   (local index i - (j+1))
    cplx* vp = v + KMAX;                            // v_{j-1} (local index i - j)
    cplx* tv = vp + KMAX;                           // t_j[c - (j+1)] = (v_j^H Q[S,S])[c]
    cplx* part = tv + KMAX;
    cplx* ring = part + RSMAX * KMAX;               // [NWARPS][RING][32]   (RING > 0 only)
    const cplx* Gb = G + (long long)b * sG;
    const cplx* tauv = tauin + (long long)b * s_vec;
    cplx* Qb = Q + (long long)b * sQ;
    const int tid = threadIdx.x, lane = tid & 31;
    for (long long x = tid; x < (long long)n * n; x += blockDim.x) Qb[x] = make_double2((x / n) == (x % n) ? 1.0 : 0.0, 0.0);
    if (n == 1) return;
    // j = n-2: S = {n-1}, t = conj(v[0]) * Q[n-1][n-1] = 1
    if (tid == 0) { tv[0] = make_double2(1.0, 0.0); }
    for (int t = tid; t < 1; t += blockDim.x) v[t] = make_double2(1.0, 0.0);
    __syncthreads();
    for (int j = n - 2; j >= 0; --j) {
        const int r0 = j + 1, ns = n - r0;
        const cplx tau = tauv[j];
        // v_j is in v[0..ns); load v_{j-1} (length ns+1) for the fused accumulation of t_{j-1}
        if (j >= 1) {
            const cplx* rowp = Gb + (long long)(j - 1) * n + j;
            for (int t = tid; t < ns + 1; t += blockDim.x) vp[t] = rowp[t];
        }
        __syncthreads();
        const Split sp = make_split(ns);
        if (sp.on) {
            const int c = r0 + sp.cb * 32 + lane;
            if constexpr (RING > 0) {
              if (c < n) {
                const cplx tc = cmul(tau, tv[c - r0]);
                cplx acc = make_double2(0.0, 0.0);
                cplx* rg = ring + ((tid >> 5) * RING) * 32 + lane;
                const int i0 = r0 + sp.rs;
                const int nr = (n - i0 + sp.rsf - 1) / sp.rsf;
                cplx* src = Qb + (long long)i0 * n + c;
                const long long step = (long long)sp.rsf * n;
#pragma unroll
                for (int q = 0; q < RING; ++q) {
                    if (q < nr) ring_load(rg + q * 32, src + q * step);
                    ring_commit();
                }
                int slot = 0;
                for (int r = 0; r < nr; ++r) {
                    ring_wait<RING - 1>();
                    cplx q = rg[slot * 32];
                    const int ii = i0 + r * sp.rsf;
                    const cplx vi = v[ii - r0];
                    q.x -= vi.x * tc.x - vi.y * tc.y;
                    q.y -= vi.x * tc.y + vi.y * tc.x;
                    src[r * step] = q;
                    if (j >= 1) acc = cfma(cconj(vp[ii - j]), q, acc);
                    if (r + RING < nr) ring_load(rg + slot * 32, src + (r + RING) * step);
                    ring_commit();
                    slot = (slot + 1 == RING) ? 0 : slot + 1;
                }
                ring_wait<0>();
                part[sp.rs * KMAX + (c - r0)] = acc;
              }
            } else if (c < n) {
                const cplx tc = cmul(tau, tv[c - r0]);
                cplx acc = make_double2(0.0, 0.0);
                int i = r0 + sp.rs;
                for (; i + 3 * sp.rsf < n; i += 4 * sp.rsf) {
                    cplx q[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) q[u] = Qb[(long long)(i + u * sp.rsf) * n + c];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int ii = i + u * sp.rsf;
                        const cplx vi = v[ii - r0];
                        q[u].x -= vi.x * tc.x - vi.y * tc.y;
                        q[u].y -= vi.x * tc.y + vi.y * tc.x;
                        Qb[(long long)ii * n + c] = q[u];
                        if (j >= 1) acc = cfma(cconj(vp[ii - j]), q[u], acc);
                    }
                }
                for (; i < n; i += sp.rsf) {
                    cplx q = Qb[(long long)i * n + c];
                    const cplx vi = v[i - r0];
                    q.x -= vi.x * tc.x - vi.y * tc.y;
                    q.y -= vi.x * tc.y + vi.y * tc.x;
                    Qb[(long long)i * n + c] = q;
                    if (j >= 1) acc = cfma(cconj(vp[i - j]), q, acc);
                }
                part[sp.rs * KMAX + (c - r0)] = acc;
            }
        }
        __syncthreads();
        if (j >= 1) {
            // t_{j-1}: local index 0 is column j (identity column/row of Q so far): conj(v_{j-1}[0]) * 1 = 1
            for (int t = tid; t < ns; t += blockDim.x) {
                cplx acc = part[t];
                for (int r = 1; r < sp.rsf; ++r) acc = cadd(acc, part[r * KMAX + t]);
                part[(RSMAX - 1) * KMAX + t] = acc;      // stash (rsf <= RSMAX; slot RSMAX-1 is read above only when rsf == RSMAX)
            }
            __syncthreads();
            for (int t = tid; t < ns + 1; t += blockDim.x) {
                tv[t] = (t == 0) ? make_double2(1.0, 0.0) : part[(RSMAX - 1) * KMAX + (t - 1)];
                v[t] = vp[t];
            }
            __syncthreads();
        }
    }
}

// Implicit QL on the real symmetric tridiagonal (d, e) with accumulation of the rotations into ZR (row c = eigenvector
// c, stored [c][component]).  Warp-specialised: thread 0 (the producer) runs the scalar QL recurrence -- a pure latency
// chain that never looks at Z -- one sweep AHEAD and leaves the rotations (c_i, s_i) of that sweep in shared memory;
// the NT consumer threads (thread 32 + t owns component t of every eigenvector) apply the previous sweep's rotations
// to Z, streaming its rows with an 8-deep register prefetch.  One barrier per sweep.
// On exit: d = eigenvalues (unsorted), Zt = ZR as complex, meta[5] = 1 if converged.
struct SweepDesc { int m, iend, valid, fail; };

template <int NT>
__global__ void __launch_bounds__(NT + 32) k_steqr(double* dio, const double* ein, long long s_vec, double* ZR, long long sZR, cplx* Zt, long long sZ,
                                                    const int* ek, int esk, int* meta, int* status, const int* mask) {
    __shared__ double d[KMAX], e[KMAX];
    __shared__ double cs[2][KMAX], sn[2][KMAX];
    __shared__ SweepDesc desc[2];
    const int b = blockIdx.x;
    if (mask && !mask[b]) return;
    const int k = ek[(long long)b * esk];
    if (k <= 0) return;
    const int tid = threadIdx.x, nthr = NT + 32;
    const int t = tid - 32;                        // component owned by a consumer thread
    double* zr = ZR + (long long)b * sZR;
    for (int i = tid; i < k; i += nthr) { d[i] = dio[(long long)b * s_vec + i]; e[i] = (i < k - 1) ? ein[(long long)b * s_vec + i] : 0.0; }
    for (long long x = tid; x < (long long)k * k; x += nthr) zr[x] = ((x / k) == (x % k)) ? 1.0 : 0.0;
    __syncthreads();
    const double eps = 2.220446049250313e-16;
    // producer state (meaningful in thread 0 only)
    int l = 0, iter = 0;
    auto produce = [&](int buf) {
        SweepDesc sd; sd.valid = 0; sd.fail = 0; sd.m = 0; sd.iend = 0;
        while (l < k) {
            int m = l;
            for (; m < k - 1; ++m) {
                const double dd = fabs(d[m]) + fabs(d[m + 1]);
                if (fabs(e[m]) <= eps * dd) break;
            }
            if (m == l) { ++l; iter = 0; continue; }
            if (++iter > 80) { sd.fail = 1; break; }
            double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
            double r = sqrt(g * g + 1.0);
            g = d[m] - d[l] + e[l] / (g + copysign(r, g));
            double s = 1.0, c = 1.0, p = 0.0;
            int i = m - 1;
            bool under = false;
            // operands of rotation i are pre-loaded one iteration ahead so that no shared-memory latency sits on the chain
            double e_i = e[i], d_i = d[i], d_ip1 = d[i + 1];
            for (; i >= l; --i) {
                double e_n = 0.0, d_n = 0.0;
                if (i - 1 >= l) { e_n = e[i - 1]; d_n = d[i - 1]; }
                const double f = s * e_i, bq = c * e_i;
                const double h2 = f * f + g * g;
                if (h2 == 0.0) { e[i + 1] = 0.0; d[i + 1] = d_ip1 - p; e[m] = 0.0; under = true; break; }
                const double ri = rsqrt(h2);
                r = h2 * ri;
                e[i + 1] = r;
                s = f * ri; c = g * ri;
                g = d_ip1 - p;
                r = (d_i - g) * s + 2.0 * c * bq;
                p = s * r;
                d[i + 1] = g + p;
                g = c * r - bq;
                cs[buf][i] = c; sn[buf][i] = s;
                e_i = e_n; d_ip1 = d_i; d_i = d_n;
            }
            if (!under) { d[l] -= p; e[l] = g; e[m] = 0.0; }
            sd.valid = 1; sd.m = m; sd.iend = under ? i + 1 : l;
            break;
        }
        desc[buf] = sd;
    };
    if (tid == 0) produce(0);
    __syncthreads();
    int cur = 0, fail = 0;
    while (true) {
        const SweepDesc sd = desc[cur];
        if (!sd.valid) { fail = sd.fail; break; }
        if (tid == 0) {
            produce(cur ^ 1);
        } else if (t >= 0 && t < k) {
            const double* cb = cs[cur];
            const double* sb = sn[cur];
            int i = sd.m - 1;
            double carry = zr[(long long)sd.m * k + t];
            double zn[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) zn[u] = (i - u >= sd.iend) ? zr[(long long)(i - u) * k + t] : 0.0;
            while (i >= sd.iend) {
                double zc[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) zc[u] = zn[u];
#pragma unroll
                for (int u = 0; u < 8; ++u) zn[u] = (i - 8 - u >= sd.iend) ? zr[(long long)(i - 8 - u) * k + t] : 0.0;
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    if (i - u >= sd.iend) {
                        const double c = cb[i - u], s = sb[i - u];
                        zr[(long long)(i - u + 1) * k + t] = s * zc[u] + c * carry;
                        carry = c * zc[u] - s * carry;
                    }
                }
                i -= 8;
            }
            zr[(long long)sd.iend * k + t] = carry;
        }
        __syncthreads();
        cur ^= 1;
    }
    __syncthreads();
    for (int i = tid; i < k; i += nthr) dio[(long long)b * s_vec + i] = d[i];
    cplx* zt = Zt + (long long)b * sZ;
    for (long long x = tid; x < (long long)k * k; x += nthr) zt[x] = make_double2(zr[x], 0.0);
    if (tid == 0) {
        if (meta) meta[b * 8 + 5] = fail ? 0 : 1;
        if (fail && status) atomicOr(&status[b], 1);
    }
}

// Divide & conquer replacement of k_steqr (yb_stedc.cuh): same contract -- d = eigenvalues (unsorted), Zt = eigenvectors as
// complex rows (row c = eigenvector c), meta[5] = 1 if converged.  ZR and the storage of Zt itself (2 k^2 doubles) are the two
// real work matrices of the merge tree.
template <int KM>
__global__ void __launch_bounds__(stedc::THREADS, 2) k_stedc(double* dio, const double* ein, long long s_vec, double* ZR, long long sZR, cplx* Zt,
                                                              long long sZ, const int* ek, int esk, int* meta, int* status, const int* mask, int expand = 1) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int b = blockIdx.x;
    if (mask && !mask[b]) return;
    const int k = ek[(long long)b * esk];
    if (k <= 0) return;
    stedc::Smem<KM>& S = *reinterpret_cast<stedc::Smem<KM>*>(smem_raw);
    double* QA = ZR + (long long)b * sZR;
    cplx* zt = Zt + (long long)b * sZ;
    double* QB = reinterpret_cast<double*>(zt);
    int fail = 0;
    double* q = stedc::solve<KM>(S, dio + (long long)b * s_vec, ein + (long long)b * s_vec, k, QA, QB, blockDim.x, &fail);
    const int tid = threadIdx.x, nthr = blockDim.x;
    const long long kk = (long long)k * k;
    // (four independent loads in flight per thread: these two passes were 8 % of the kernel on one dependent load each)
    if (q != QA) {
        long long x = tid;
        for (; x + 3LL * nthr < kk; x += 4LL * nthr) {
            const double a0 = QB[x], a1 = QB[x + nthr], a2 = QB[x + 2LL * nthr], a3 = QB[x + 3LL * nthr];
            QA[x] = a0; QA[x + nthr] = a1; QA[x + 2LL * nthr] = a2; QA[x + 3LL * nthr] = a3;
        }
        for (; x < kk; x += nthr) QA[x] = QB[x];
        __syncthreads();
    }
    if (expand) {                        // (the "top" mode of yb_heev_rows_top gathers the rows it needs from QA itself)
        long long x = tid;
        for (; x + 3LL * nthr < kk; x += 4LL * nthr) {
            const double a0 = QA[x], a1 = QA[x + nthr], a2 = QA[x + 2LL * nthr], a3 = QA[x + 3LL * nthr];
            zt[x] = make_double2(a0, 0.0); zt[x + nthr] = make_double2(a1, 0.0);
            zt[x + 2LL * nthr] = make_double2(a2, 0.0); zt[x + 3LL * nthr] = make_double2(a3, 0.0);
        }
        for (; x < kk; x += nthr) zt[x] = make_double2(QA[x], 0.0);
    }
    if (tid == 0) {
        if (meta) meta[b * 8 + 5] = fail ? 0 : 1;
        if (fail && status) atomicOr(&status[b], 1);
    }
}

// Descending order of the eigenvalues of one trajectory: order[rank] = index, dsorted[rank] = value, ntop = min(k, top).
__global__ void __launch_bounds__(512) k_heev_order(const double* __restrict__ d, long long s_vec, const int* ek, int esk, int top, int* order,
                                                     double* dsorted, int* ntop, const int* mask) {
    __shared__ double ev[KMAX];
    const int b = blockIdx.x;
    if (mask && !mask[b]) return;
    const int k = ek[(long long)b * esk];
    if (k <= 0) { if (threadIdx.x == 0) ntop[b] = 0; return; }
    const double* db = d + (long long)b * s_vec;
    for (int i = threadIdx.x; i < k; i += blockDim.x) ev[i] = db[i];
    __syncthreads();
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        const double vi = ev[i];
        int rank = 0;
        for (int j = 0; j < k; ++j) { const double vj = ev[j]; rank += (vj > vi) || (vj == vi && j < i) || (vj == vj && vi != vi); }     // NaN last
        order[(long long)b * s_vec + rank] = i;
        dsorted[(long long)b * s_vec + rank] = vi;
    }
    if (threadIdx.x == 0) ntop[b] = top > 0 ? min(k, top) : k;
}

}  // namespace heev
