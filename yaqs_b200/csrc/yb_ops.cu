// Memory-bound batched kernels of the TJM hot path: MPO-tensor application, Lanczos vector work + the tiny
// tridiagonal eigen-solve, site-tensor utilities, local reduced density matrices.  All are HBM-bound streaming
// kernels: coalesced 16-byte accesses along the fastest (right-bond) index, one launch covers the whole batch.
#include "yb_common.cuh"
#include "yb_philox.cuh"
#include "yb_mpo_apply.cuh"
#include "yb_wapply.cuh"
#include <math.h>

// ---------------------------------------------------------------------------------------------------------------------
// out[x,l,y,r] = sum_(u,v) wt[x,u,y,v] in[u,l,v,r]      (primitives.py:106,135,202)
// grid = (ceil(l*r / 256), nx*ny, B).  wt is tiny and identical for all trajectories -> read through the read-only path;
// zero entries of the (sparse, ~30 % dense) MPO tensor are skipped block-uniformly.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void k_wapply(int nx, int nu, int ny, int nv, const cplx* __restrict__ wt, const cplx* __restrict__ in, long long s_in,
                         cplx* __restrict__ out, long long s_out, const int* el, int esl, const int* er, int esr, const int* mask) {
    const int b = blockIdx.z;
    if (mask && !mask[b]) return;
    const int l = el[(long long)b * esl], r = er[(long long)b * esr];
    const long long lr = (long long)l * r;
    const long long pos = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= lr) return;
    const int x = blockIdx.y / ny, y = blockIdx.y % ny;
    const int li = (int)(pos / r), ri = (int)(pos % r);
    const cplx* src = in + (long long)b * s_in;
    cplx acc = make_double2(0.0, 0.0);
    for (int u = 0; u < nu; ++u)
        for (int v = 0; v < nv; ++v) {
            const cplx w = __ldg(&wt[((x * nu + u) * ny + y) * nv + v]);
            if (w.x == 0.0 && w.y == 0.0) continue;
            const cplx t = src[(((long long)u * l + li) * nv + v) * r + ri];
            acc = cfma(w, t, acc);
        }
    out[(long long)b * s_out + (((long long)x * l + li) * ny + y) * r + ri] = acc;
}

// One-pass variant (yb_wapply.cuh): a thread owns one (l, r) position and produces all nx * ny outputs from one read of the input.
// grid = (ceil(l*r / 128), 1, B); wt staged in shared memory (<= 400 complex).
__global__ void __launch_bounds__(128) k_wapply_onepass(int nx, int nu, int ny, int nv, const cplx* __restrict__ wt, const cplx* __restrict__ in,
                                                        long long s_in, cplx* __restrict__ out, long long s_out, const int* el, int esl,
                                                        const int* er, int esr, const int* mask) {
    __shared__ cplx ws[wapply::MAX_NX * wapply::MAX_NX * wapply::MAX_NY * wapply::MAX_NY];
    const int b = blockIdx.z;
    if (mask && !mask[b]) return;
    const int l = el[(long long)b * esl], r = er[(long long)b * esr];
    const long long lr = (long long)l * r;
    if ((long long)blockIdx.x * blockDim.x >= lr) return;                 // block-uniform: nobody waits at the barrier below
    for (int e = threadIdx.x; e < nx * nu * ny * nv; e += blockDim.x) ws[e] = wt[e];
    __syncthreads();
    const long long pos = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= lr) return;
    wapply::element(pos, nx, nu, ny, nv, l, r, ws, in + (long long)b * s_in, out + (long long)b * s_out);
}

// env[i,a,j] = delta_ij
__global__ void k_env_identity(int dmpo, cplx* env, long long s_env, const int* echi, int es, const int* mask) {
    const int b = blockIdx.y;
    if (mask && !mask[b]) return;
    const int chi = echi[(long long)b * es];
    const long long n = (long long)chi * dmpo * chi;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(e % chi);
        const int i = (int)(e / ((long long)dmpo * chi));
        env[(long long)b * s_env + e] = make_double2(i == j ? 1.0 : 0.0, 0.0);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Lanczos (matrix_exponential.py:33-201).  One CTA per trajectory.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int LZ_THREADS = 1024;
constexpr int LZ_KMAX = 32;

__global__ void __launch_bounds__(LZ_THREADS) k_lanczos_init(int dphys, const cplx* __restrict__ x, long long s_x, cplx* basis, long long s_basis,
                                                               double* nrm, int* active, int* kdone, const int* el, int esl, const int* er,
                                                               int esr, const int* mask, int* n_zero) {
    __shared__ double red[32];
    const int b = blockIdx.x;
    if (n_zero && b == 0 && threadIdx.x == 0) *n_zero = 0;      // counter of the first Lanczos step (no separate fill launch)
    if (mask && !mask[b]) { active[b] = 0; kdone[b] = -1; return; }
    const long long n = (long long)dphys * el[(long long)b * esl] * er[(long long)b * esr];
    const cplx* xb = x + (long long)b * s_x;
    cplx* v0 = basis + (long long)b * s_basis;
    double s[1] = {0.0};
    for (long long e = threadIdx.x; e < n; e += blockDim.x) s[0] += cabs2(xb[e]);
    block_sum<1>(s, red);
    const double nr = sqrt(s[0]);
    if (threadIdx.x == 0) { nrm[b] = nr; active[b] = nr > 0.0 ? 1 : 0; kdone[b] = 0; }
    if (nr == 0.0) return;                                   // zero vector: returned untouched (:67-69)
    const double inv = 1.0 / nr;
    for (long long e = threadIdx.x; e < n; e += blockDim.x) { const cplx t = xb[e]; v0[e] = make_double2(t.x * inv, t.y * inv); }
}

// Symmetric tridiagonal eigen-decomposition by implicit QL (one warp; lane r owns row r of Z).
// d[0..k) diagonal -> eigenvalues, e[0..k-1) off-diagonal (destroyed), z row-major [k][LZ_KMAX].
__device__ void tridiag_eig_warp(int k, double* d, double* e, double* z, int lane) {
    for (int i = lane; i < k * LZ_KMAX; i += 32) z[i] = 0.0;
    __syncwarp();
    if (lane < k) z[lane * LZ_KMAX + lane] = 1.0;
    if (lane == 0) e[k - 1] = 0.0;
    __syncwarp();
    const double eps = 2.220446049250313e-16;
    for (int l = 0; l < k; ++l) {
        for (int iter = 0; iter < 60; ++iter) {
            int m = l;
            for (; m < k - 1; ++m) {
                const double dd = fabs(d[m]) + fabs(d[m + 1]);
                if (fabs(e[m]) <= eps * dd) break;
            }
            if (m == l) break;
            double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
            double r = hypot(g, 1.0);
            g = d[m] - d[l] + e[l] / (g + copysign(r, g));
            double s = 1.0, c = 1.0, p = 0.0;
            int i = m - 1;
            bool underflow = false;
            for (; i >= l; --i) {
                double f = s * e[i];
                const double bb = c * e[i];
                r = hypot(f, g);
                __syncwarp();
                if (lane == 0) e[i + 1] = r;
                if (r == 0.0) {
                    if (lane == 0) { d[i + 1] -= p; e[m] = 0.0; }
                    underflow = true;
                    break;
                }
                s = f / r; c = g / r;
                g = d[i + 1] - p;
                r = (d[i] - g) * s + 2.0 * c * bb;
                p = s * r;
                __syncwarp();
                if (lane == 0) d[i + 1] = g + p;
                g = c * r - bb;
                if (lane < k) {
                    const double f2 = z[lane * LZ_KMAX + i + 1], zi = z[lane * LZ_KMAX + i];
                    z[lane * LZ_KMAX + i + 1] = s * zi + c * f2;
                    z[lane * LZ_KMAX + i] = c * zi - s * f2;
                }
                __syncwarp();
            }
            __syncwarp();
            if (underflow) continue;
            if (lane == 0) { d[l] -= p; e[l] = g; e[m] = 0.0; }
            __syncwarp();
        }
    }
    __syncwarp();
}

__global__ void __launch_bounds__(LZ_THREADS) k_lanczos_step(int j, int kmax, double dt, double tol, int dphys, cplx* basis, long long s_basis,
                                                               long long s_vec, double* alpha, double* beta, const double* nrm, int* active,
                                                               int* kdone, cplx* coef, int* n_active, const int* el, int esl, const int* er,
                                                               int esr, int* n_zero) {
    __shared__ double red[64];
    __shared__ double sd[LZ_KMAX], se[LZ_KMAX], sz[LZ_KMAX * LZ_KMAX];
    __shared__ int s_finish, s_k;
    const int b = blockIdx.x;
    if (n_zero && b == 0 && threadIdx.x == 0) *n_zero = 0;      // the OTHER counter of the pair: the next step counts into it
    if (!active[b]) return;
    const long long n = (long long)dphys * el[(long long)b * esl] * er[(long long)b * esr];
    cplx* vj = basis + (long long)b * s_basis + (long long)j * s_vec;
    cplx* w = vj + s_vec;
    const cplx* vp = vj - s_vec;
    double* al = alpha + (long long)b * LZ_KMAX;
    double* be = beta + (long long)b * LZ_KMAX;
    // alpha_j = Re <v_j, w>   (:117)
    double s[1] = {0.0};
    for (long long e = threadIdx.x; e < n; e += blockDim.x) { const cplx a = vj[e], c = w[e]; s[0] += a.x * c.x + a.y * c.y; }
    block_sum<1>(s, red);
    const double aj = s[0];
    const double bprev = j > 0 ? be[j - 1] : 0.0;
    // w -= alpha_j v_j + beta_{j-1} v_{j-1}; beta_j = |w|   (:121-128)
    double t[1] = {0.0};
    for (long long e = threadIdx.x; e < n; e += blockDim.x) {
        cplx c = w[e];
        const cplx a = vj[e];
        c.x -= aj * a.x; c.y -= aj * a.y;
        if (j > 0) { const cplx p = vp[e]; c.x -= bprev * p.x; c.y -= bprev * p.y; }
        w[e] = c;
        t[0] += cabs2(c);
    }
    block_sum<1>(t, red);
    const double bj = sqrt(t[0]);
    const bool last = (j == kmax - 1);
    const double eps_cut = 100.0 * (double)n * 2.220446049250313e-16;
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        if (lane == 0) { al[j] = aj; if (!last) be[j] = bj; }
        __syncwarp();
        int finish = 0, k = j + 1;
        const bool breakdown = (!last) && (bj < eps_cut);
        if (breakdown || j >= 1 || last) {
            if (k == 1) {
                if (lane == 0) { sd[0] = aj; sz[0] = 1.0; }
            } else {
                for (int i = lane; i < k; i += 32) { sd[i] = al[i]; se[i] = (i < k - 1) ? be[i] : 0.0; }
                __syncwarp();
                tridiag_eig_warp(k, sd, se, sz, lane);
            }
            __syncwarp();
            if (breakdown || last) finish = 1;
            else {
                // err = beta_j * | sum_m z[k-1,m] exp(-i dt lam_m) z[0,m] |   (:154-160)
                double pr = 0.0, pi = 0.0;
                for (int m = 0; m < k; ++m) {
                    double sn, cs; sincos(-dt * sd[m], &sn, &cs);
                    const double ww = sz[(k - 1) * LZ_KMAX + m] * sz[m];
                    pr += ww * cs; pi += ww * sn;
                }
                if (bj * sqrt(pr * pr + pi * pi) < tol) finish = 1;
            }
            if (finish && lane < k) {
                // y_i = |v| sum_m z[i,m] exp(-i dt lam_m) z[0,m]   (:162-163, :199-201)
                double pr = 0.0, pi = 0.0;
                for (int m = 0; m < k; ++m) {
                    double sn, cs; sincos(-dt * sd[m], &sn, &cs);
                    const double ww = sz[lane * LZ_KMAX + m] * sz[m];
                    pr += ww * cs; pi += ww * sn;
                }
                const double nr = nrm[b];
                coef[(long long)b * LZ_KMAX + lane] = make_double2(nr * pr, nr * pi);
            }
        }
        if (lane == 0) {
            s_finish = finish; s_k = k;
            if (finish) { kdone[b] = k; active[b] = 0; } else atomicAdd(n_active, 1);
        }
    }
    __syncthreads();
    if (s_finish) return;
    // v_{j+1} = w / beta_j   (:134)
    const double inv = 1.0 / bj;
    for (long long e = threadIdx.x; e < n; e += blockDim.x) { cplx c = w[e]; c.x *= inv; c.y *= inv; w[e] = c; }
}

__global__ void k_lanczos_combine(int dphys, const cplx* __restrict__ basis, long long s_basis, long long s_vec, const cplx* __restrict__ coef,
                                  const int* kdone, cplx* out, long long s_out, const int* el, int esl, const int* er, int esr, const int* mask) {
    const int b = blockIdx.y;
    if (mask && !mask[b]) return;
    const int k = kdone[b];
    if (k <= 0) return;
    const long long n = (long long)dphys * el[(long long)b * esl] * er[(long long)b * esr];
    __shared__ cplx cf[LZ_KMAX];
    if (threadIdx.x < k) cf[threadIdx.x] = coef[(long long)b * LZ_KMAX + threadIdx.x];
    __syncthreads();
    const cplx* vb = basis + (long long)b * s_basis;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        cplx acc = make_double2(0.0, 0.0);
        for (int i = 0; i < k; ++i) acc = cfma(cf[i], vb[(long long)i * s_vec + e], acc);
        out[(long long)b * s_out + e] = acc;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// site-tensor utilities
// ---------------------------------------------------------------------------------------------------------------------
// dst(s, k, r) = src(k, (s, r))
// set_ext (optional): after the copy the trajectory's extent at set_ext := k (the new bond dimension; no launch reads it meanwhile)
__global__ void k_unpack_vh(int dphys, const cplx* __restrict__ src, long long s_src, cplx* dst, long long s_dst, const int* ek, int esk,
                            const int* er, int esr, const int* mask, int* set_ext, int set_stride) {
    const int b = blockIdx.y;
    if (mask && !mask[b]) return;
    const int k = ek[(long long)b * esk], r = er[(long long)b * esr];
    if (set_ext && blockIdx.x == 0 && threadIdx.x == 0) set_ext[(long long)b * set_stride] = k;
    const long long n = (long long)dphys * k * r;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const int ri = (int)(e % r);
        const long long q = e / r;
        const int ki = (int)(q % k), s = (int)(q / k);
        dst[(long long)b * s_dst + e] = src[(long long)b * s_src + ((long long)ki * dphys + s) * r + ri];
    }
}

__global__ void k_copy_packed(int mult, const cplx* __restrict__ src, long long s_src, cplx* dst, long long s_dst, const int* e0, int es0,
                              const int* e1, int es1, const int* mask, int* set_ext, int set_stride) {
    const int b = blockIdx.y;
    if (mask && !mask[b]) return;
    const long long n = (long long)mult * e0[(long long)b * es0] * (e1 ? e1[(long long)b * es1] : 1);
    if (set_ext && blockIdx.x == 0 && threadIdx.x == 0) set_ext[(long long)b * set_stride] = e0[(long long)b * es0];      // see k_unpack_vh
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x)
        dst[(long long)b * s_dst + e] = src[(long long)b * s_src + e];
}

// A_s <- sum_t op[s,t] A_t   (dphys <= 4)
__global__ void k_apply_local_op(int dphys, const cplx* __restrict__ op, long long s_op, cplx* a, long long s_a, const int* el, int esl,
                                 const int* er, int esr, const int* mask, const int* sel, int sel_value) {
    const int b = blockIdx.y;
    if (mask && !mask[b]) return;
    if (sel && sel[b] != sel_value) return;
    const long long lr = (long long)el[(long long)b * esl] * er[(long long)b * esr];
    const cplx* o = op + (long long)b * s_op;
    cplx* ab = a + (long long)b * s_a;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < lr; e += (long long)gridDim.x * blockDim.x) {
        cplx v[4], res[4];
        for (int t = 0; t < dphys; ++t) v[t] = ab[(long long)t * lr + e];
        for (int s = 0; s < dphys; ++s) {
            cplx acc = make_double2(0.0, 0.0);
            for (int t = 0; t < dphys; ++t) acc = cfma(o[s * dphys + t], v[t], acc);
            res[s] = acc;
        }
        for (int s = 0; s < dphys; ++s) ab[(long long)s * lr + e] = res[s];
    }
}

// A'[s, (l a), (r b)] = sum_t W[s, t, a, b] A[t, l, r]: one MPO site tensor applied to one MPS site tensor with the bond
// indices fused (MPS bond major).  One thread per output matrix element (l a, r b) (yb_mpo_apply.cuh); reads are gathers from a
// tensor that is Dl * Dr times smaller than the output, writes are coalesced.  dphys <= 4.
__global__ void k_mpo_site_apply(int dphys, int dl, int dr, const cplx* __restrict__ w, const cplx* __restrict__ src, long long s_src, cplx* dst,
                                 long long s_dst, const int* el, int esl, const int* er, int esr, const int* mask) {
    const int b = blockIdx.y;
    if (mask && !mask[b]) return;
    const long long l = el[(long long)b * esl], r = er[(long long)b * esr];
    const long long out_lr = l * dl * r * dr;
    const cplx* a = src + (long long)b * s_src;
    cplx* o = dst + (long long)b * s_dst;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < out_lr; e += (long long)gridDim.x * blockDim.x)
        mpo_apply::element(e, dphys, dl, dr, l, r, w, a, o);
}

__global__ void __launch_bounds__(1024) k_norm2(int dphys, const cplx* __restrict__ a, long long s_a, double* out, const int* el, int esl,
                                                 const int* er, int esr, const int* mask) {
    __shared__ double red[32];
    const int b = blockIdx.x;
    if (mask && !mask[b]) return;
    const long long n = (long long)dphys * el[(long long)b * esl] * er[(long long)b * esr];
    const cplx* ab = a + (long long)b * s_a;
    double s[1] = {0.0};
    for (long long e = threadIdx.x; e < n; e += blockDim.x) s[0] += cabs2(ab[e]);
    block_sum<1>(s, red);
    if (threadIdx.x == 0) out[b] = s[0];
}

__global__ void k_scale(int dphys, const double* scale, cplx* a, long long s_a, const int* el, int esl, const int* er, int esr, const int* mask,
                        int rsqrt_mode) {
    const int b = blockIdx.y;
    if (mask && !mask[b]) return;
    const long long n = (long long)dphys * el[(long long)b * esl] * er[(long long)b * esr];
    const double f = rsqrt_mode ? 1.0 / sqrt(scale[b]) : scale[b];       // rsqrt_mode: scale holds a squared norm (normalisation in one launch)
    cplx* ab = a + (long long)b * s_a;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        cplx v = ab[e]; v.x *= f; v.y *= f; ab[e] = v;
    }
}

// g[b][o][p] = sum_(l,r) conj(a_o[l,r]) x_p[l,r]; one CTA per trajectory, dphys <= 4.
__global__ void __launch_bounds__(512) k_local_rdm(int dphys, const cplx* __restrict__ a, long long s_a, const cplx* __restrict__ x, long long s_x,
                                                    cplx* g, const int* el, int esl, const int* er, int esr, const int* mask) {
    __shared__ double red[32 * 32];
    const int b = blockIdx.x;
    if (mask && !mask[b]) return;
    const long long lr = (long long)el[(long long)b * esl] * er[(long long)b * esr];
    const cplx* ab = a + (long long)b * s_a;
    const cplx* xb = x + (long long)b * s_x;
    double acc[32];
#pragma unroll
    for (int q = 0; q < 32; ++q) acc[q] = 0.0;
    for (long long e = threadIdx.x; e < lr; e += blockDim.x) {
        cplx av[4], xv[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) if (t < dphys) { av[t] = ab[(long long)t * lr + e]; xv[t] = xb[(long long)t * lr + e]; }
#pragma unroll
        for (int o = 0; o < 4; ++o)
#pragma unroll
            for (int p = 0; p < 4; ++p)
                if (o < dphys && p < dphys) {
                    acc[(o * 4 + p) * 2] += av[o].x * xv[p].x + av[o].y * xv[p].y;
                    acc[(o * 4 + p) * 2 + 1] += av[o].x * xv[p].y - av[o].y * xv[p].x;
                }
    }
    block_sum<32>(acc, red);
    if (threadIdx.x < dphys * dphys) {
        const int o = threadIdx.x / dphys, p = threadIdx.x % dphys;
        g[(long long)b * dphys * dphys + threadIdx.x] = make_double2(acc[(o * 4 + p) * 2], acc[(o * 4 + p) * 2 + 1]);
    }
}

// out[b][2c], out[b][2c + 1] = the two uniforms of Philox block (block0 + c) of stream (seed, traj[b], tag); one thread per
// block.  HBM-bound (16 B written per 10-round block), launched once per run.
__global__ void k_philox_uniform(unsigned long long seed, const long long* __restrict__ traj, unsigned tag, int nblocks, unsigned block0, double* out,
                                 long long s_out) {
    const int b = blockIdx.y;
    const long long t = traj[b];
    double* ob = out + (long long)b * s_out;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < nblocks; c += gridDim.x * blockDim.x) {
        double u0, u1;
        philox::pair(seed, t, tag, block0 + (unsigned)c, u0, u1);
        reinterpret_cast<double2*>(ob)[c] = make_double2(u0, u1);
    }
}


// ---------------------------------------------------------------------------------------------------------------------
// stochastic_process on the device (methods/stochastic_process.py:190-292): the jump test and the selection of the jump process.
// ---------------------------------------------------------------------------------------------------------------------
// Jump test for every trajectory: dp = 1 - <psi|psi> (calculate_stochastic_factor, :36-60), jump iff u1 < dp (:238).  The two
// uniforms a test can consume sit at positions p, p + 1 of the trajectory's stream, p = pos[b] (or `col` for all when pos == NULL):
//   table mode   u = table[b * stride + p]                       (streams replayed from the reference's seeded numpy generators)
//   philox mode  u = component (q & 1) of block (q >> 1), q = 2 * block0 + p, of stream (seed, traj[b], tag), drawn HERE
// pos[b] advances by 1 + jumped -- the doubles the reference's Generator hands out (rng.random() at :238, rng.choice at :257).
// One CTA walks the whole batch, so the number of jumps needs no atomics and no cleared counter.
__global__ void __launch_bounds__(256) k_jump_test(int nbatch, const double* __restrict__ norm2, const double* __restrict__ table, long long stride,
                                                    long long col, long long* pos, int use_philox, unsigned long long seed,
                                                    const long long* __restrict__ traj, unsigned tag, unsigned block0, double* dp_out,
                                                    double* u1_out, double* u2_out, int* jmask, int* njump) {
    __shared__ int cnt[8];
    int mine = 0;
    for (int b = threadIdx.x; b < nbatch; b += blockDim.x) {
        const long long p = pos ? pos[b] : col;
        double u1, u2;
        if (use_philox) {
            const unsigned long long q = 2ull * block0 + (unsigned long long)p;
            double a0, a1;
            philox::pair(seed, traj[b], tag, (unsigned)(q >> 1), a0, a1);
            if (q & 1) { u1 = a1; double c0, c1; philox::pair(seed, traj[b], tag, (unsigned)(q >> 1) + 1u, c0, c1); u2 = c0; }
            else { u1 = a0; u2 = a1; }
        } else {
            u1 = table[(long long)b * stride + p];
            u2 = table[(long long)b * stride + p + 1];
        }
        const double dp = 1.0 - norm2[b];
        const int jump = (u1 < dp) ? 1 : 0;
        dp_out[b] = dp; u1_out[b] = u1; u2_out[b] = u2; jmask[b] = jump;
        if (pos) pos[b] = p + 1 + jump;
        mine += jump;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if ((threadIdx.x & 31) == 0) cnt[threadIdx.x >> 5] = mine;
    __syncthreads();
    if (threadIdx.x == 0) { int s = 0; for (int w = 0; w < 8; ++w) s += cnt[w]; *njump = s; }
}

// Selection of the jump process for the trajectories that jump (create_probability_distribution + rng.choice, :86-187, :257):
//   w_k = dt gamma_k Re Tr(L_k^dagger L_k rho_site(k))   from the local reduced density matrices of the transfer sweep
//   total = sum_k w_k (zero / non-finite: status bit 4, :179-187), pdf = w / total, cdf = cumsum(pdf) / cdf[-1],
//   choice = #{k : cdf_k <= u2} (numpy's Generator.choice: searchsorted(cdf, u, side="right")), clamped to the last process.
// All sums run in process order, one thread per trajectory.  ops_sel[b] = ops_table[choice]: the operator(s) the application launches take.
__global__ void k_jump_select(int nbatch, const int* __restrict__ jmask, const double* __restrict__ u2, const cplx* __restrict__ rdm1,
                              const cplx* __restrict__ rdm2, int nproc, const int* __restrict__ psite, const int* __restrict__ ptwo,
                              const cplx* __restrict__ pldl, const double* __restrict__ pstrength, double dt, const cplx* __restrict__ ops_table,
                              int nops, double* w, double* total_out, int* choice_out, cplx* ops_sel, int* status) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nbatch) return;
    if (!jmask[b]) { total_out[b] = 0.0; choice_out[b] = -1; return; }
    double* wb = w + (long long)b * nproc;
    double total = 0.0;
    for (int k = 0; k < nproc; ++k) {
        const int site = psite[k];
        const cplx* m = pldl + (long long)k * 16;
        double acc = 0.0;
        if (ptwo[k]) {
            const cplx* g = rdm2 + ((long long)site * nbatch + b) * 16;
            for (int e = 0; e < 16; ++e) acc += m[e].x * g[e].x - m[e].y * g[e].y;
        } else {
            const cplx* g = rdm1 + ((long long)site * nbatch + b) * 4;
            for (int e = 0; e < 4; ++e) acc += m[e].x * g[e].x - m[e].y * g[e].y;
        }
        const double wk = acc * (dt * pstrength[k]);
        wb[k] = wk;
        total += wk;
    }
    total_out[b] = total;
    const bool ok = isfinite(total) && total > 0.0;
    if (!ok) atomicOr(&status[b], 4);
    const double safe = ok ? total : 1.0;
    double last = 0.0;
    for (int k = 0; k < nproc; ++k) { const double p = wb[k] / safe; wb[k] = p; last += p; }
    int choice = 0;
    double cum = 0.0;
    for (int k = 0; k < nproc; ++k) { cum += wb[k]; choice += ((cum / last) <= u2[b]) ? 1 : 0; }
    choice = min(choice, nproc - 1);
    choice_out[b] = choice;
    for (int e = 0; e < nops; ++e) ops_sel[(long long)b * nops + e] = ops_table[(long long)choice * nops + e];
}


// Expectation values of the local observables from the reduced density matrices of the transfer sweep (mps.py:1178-1234):
// out[b][col_k] = Re sum_e mats[k][e] rdm[site_k][b][e]; an imaginary part >= 1e-13 raises status bit 8 (the reference asserts, :1233).
__global__ void k_observables(int nbatch, int nobs, const int* __restrict__ col, const int* __restrict__ site, const int* __restrict__ two,
                              const cplx* __restrict__ mats, const cplx* __restrict__ rdm1, const cplx* __restrict__ rdm2, double* out,
                              long long out_stride, int* status) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)nbatch * nobs) return;
    const int k = (int)(t / nbatch), b = (int)(t - (long long)k * nbatch);       // consecutive threads = consecutive trajectories of one observable
    const cplx* m = mats + (long long)k * 16;
    const int ne = two[k] ? 16 : 4;
    const cplx* g = two[k] ? rdm2 + ((long long)site[k] * nbatch + b) * 16 : rdm1 + ((long long)site[k] * nbatch + b) * 4;
    double re = 0.0, im = 0.0;
    for (int e = 0; e < ne; ++e) { re += m[e].x * g[e].x - m[e].y * g[e].y; im += m[e].x * g[e].y + m[e].y * g[e].x; }
    out[(long long)b * out_stride + col[k]] = re;
    if (im >= 1e-13) atomicOr(&status[b], 8);
}

// Bond diagnostics of every trajectory (mps.py:549-602): out[b] = (sum chi^3, max(d, max chi), sum chi) over the inner bonds 1 .. L-1.
__global__ void k_diagnostics(int nbatch, int length, int dphys, const int* __restrict__ bond, double* out, long long out_stride) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nbatch) return;
    const int* bb = bond + (long long)b * (length + 1);
    double cost = 0.0, total = 0.0;
    int mx = dphys;
    for (int j = 1; j < length; ++j) { const double c = (double)bb[j]; cost += c * c * c; total += c; mx = max(mx, bb[j]); }
    out[(long long)b * out_stride + 0] = cost;
    out[(long long)b * out_stride + 1] = (double)mx;
    out[(long long)b * out_stride + 2] = total;
}


// work[b] = (mask == NULL || mask[b]) && !flag[b]: the trajectories of a masked call that a certificate accepted (the launches that
// install the factors of a Cholesky-QR shift run under it; the flagged rest is redone by the eigen-solver route and installed afterwards)
__global__ void k_mask_andnot(int nbatch, const int* __restrict__ mask, const int* __restrict__ flag, int* work) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < nbatch) work[b] = ((!mask || mask[b]) && !flag[b]) ? 1 : 0;
}
