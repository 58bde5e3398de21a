// Blocked Hermitian tridiagonalisation and the matching back-transformation (one CTA per trajectory) -- round 2 replacement of the
// unblocked k_hetrd / k_ungtr pair of yb_heev.cu on the Gram-matrix SVD route (replaces the bulk of scipy.linalg.svd / zgesdd at
// core/linalg/svd.py:89-104 for the matrices of split_two_site, core/methods/decompositions.py:105-185).
//
// Why: the unblocked zhetd2 reads AND writes the whole trailing block once per Householder vector -- 179 MB of L2/HBM traffic per
// 256 x 256 matrix, 110 MB of it DRAM (profiles/r01_ncu_heev_full_stedc.json), 4.8 ms for a wave of 148 matrices -- and k_ungtr
// streamed a second k x k matrix the same way only to feed one more GEMM.  Here:
//   k_hetrd_blocked  zlatrd panels of NB columns: inside a panel the trailing block is only READ (one Hermitian matrix-vector
//                    product per column; the rank-2 corrections of the columns already reduced live in two n x NB panels V, W in
//                    shared memory), then ONE rank-2NB update A -= V W^H + W V^H of the trailing block per panel, computed on its
//                    lower triangle in 32 x 32 tiles (4-row register strips) and mirrored.  Writes drop from n^3/3 to n^3/(3 NB)
//                    elements, reads stay n^3/3: 179 MB -> ~95 MB per 256 x 256 matrix.
//   k_ormtr_rows     the eigenvector rows E^H = Z^T Qh^H straight from the reflectors: every warp owns rows of Z^T and applies the
//                    reflectors of a shared-memory panel one after the other (dot + axpy on registers).  No explicit Qh, no k^3
//                    GEMM afterwards (the EVEC product of round 1).
// Conventions are those of yb_heev.cu: G row-major, full Hermitian storage; on exit row j of G holds v_j (v_j[0] = 1 at column j+1,
// zeros up to column j), H_j = I - tau_j v_j v_j^H, Qh = H_0 ... H_{n-2}, G = Qh T Qh^H with T = tridiag(d, e).  The algebra was
// checked against numpy on the host before this kernel was written (tests/test_hetrd_model.py restates it).
#pragma once

namespace hetrd {

constexpr int THREADS = 512;
constexpr int NWARPS = THREADS / 32;
constexpr int RSMAX = 4;                        // row groups per column block of the matrix-vector product

// shared memory (complex elements): V, W panels [nmax][NB + 1]; vcur, acol/p [nmax]; part [RSMAX][nmax]; small vectors
template <int NB>
constexpr size_t smem_bytes(int nmax) { return ((size_t)(2 * (NB + 1) + 2 + RSMAX) * nmax + 4 * NB) * sizeof(cplx) + 64 * sizeof(double); }

template <int NB>
__global__ void __launch_bounds__(THREADS, 1) k_hetrd_blocked(cplx* G, long long sG, const int* ek, int esk, double* dout, double* eout, cplx* tauout,
                                                               long long s_vec, int nmax, const int* mask) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int PL = NB + 1;                 // panel pitch: odd multiple of 16 bytes -> conflict-free lane-per-row reads
    const int b = blockIdx.x;
    if (mask && !mask[b]) return;
    const int n = ek[(long long)b * esk];
    if (n <= 0) return;
    cplx* Vs = reinterpret_cast<cplx*>(smem_raw);
    cplx* Ws = Vs + (size_t)nmax * PL;
    cplx* vcur = Ws + (size_t)nmax * PL;
    cplx* pcur = vcur + nmax;
    cplx* part = pcur + nmax;                  // [RSMAX][nmax]
    cplx* tw = part + (size_t)RSMAX * nmax;    // [NB]  W^H v
    cplx* tv = tw + NB;                        // [NB]  V^H v
    cplx* wj = tv + NB;                        // [NB]  conj(W[j][:])
    cplx* vj = wj + NB;                        // [NB]  conj(V[j][:])
    double* red = reinterpret_cast<double*>(vj + NB);
    cplx* Gb = G + (long long)b * sG;
    double* d = dout + (long long)b * s_vec;
    double* e = eout + (long long)b * s_vec;
    cplx* tauv = tauout + (long long)b * s_vec;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (n == 1) {
        if (tid == 0) { d[0] = Gb[0].x; e[0] = 0.0; tauv[0] = make_double2(0.0, 0.0); Gb[0] = make_double2(0.0, 0.0); }
        return;
    }
    for (int j0 = 0; j0 < n - 1; j0 += NB) {
        const int nb = min(NB, n - 1 - j0);
        for (int i = 0; i < nb; ++i) {
            const int j = j0 + i, ns = n - 1 - j;          // S = {j+1 .. n-1}
            // ---- (1) column j of the current matrix, rows j .. n-1:  a = conj(G[j][j:]) - V conj(W[j]) - W conj(V[j]) -----------
            if (tid < i) { wj[tid] = cconj(Ws[(size_t)j * PL + tid]); vj[tid] = cconj(Vs[(size_t)j * PL + tid]); }
            __syncthreads();
            for (int t = tid; t <= ns; t += THREADS) {
                const int r = j + t;
                cplx a = cconj(Gb[(long long)j * n + r]);
                const cplx* vr = Vs + (size_t)r * PL;
                const cplx* wr = Ws + (size_t)r * PL;
                for (int q = 0; q < i; ++q) {
                    const cplx x = cmul(vr[q], wj[q]), y = cmul(wr[q], vj[q]);
                    a.x -= x.x + y.x; a.y -= x.y + y.y;
                }
                pcur[t] = a;
            }
            __syncthreads();
            // ---- (2) zlarfg on x = a[1:] (alpha = a[1]): v (v[0] = 1), tau, beta ------------------------------------------------
            double s2[1] = {0.0};
            for (int t = 2 + tid; t <= ns; t += THREADS) s2[0] += cabs2(pcur[t]);
            block_sum<1>(s2, red);
            const cplx alpha = pcur[1];
            const double dj = pcur[0].x;
            cplx tau; double beta;
            if (s2[0] == 0.0 && alpha.y == 0.0) {
                tau = make_double2(0.0, 0.0); beta = alpha.x;
                for (int t = tid; t < ns; t += THREADS) vcur[t] = make_double2(t == 0 ? 1.0 : 0.0, 0.0);
            } else {
                beta = -copysign(sqrt(alpha.x * alpha.x + alpha.y * alpha.y + s2[0]), alpha.x);
                tau = make_double2((beta - alpha.x) / beta, -alpha.y / beta);
                const double dr = alpha.x - beta, di = alpha.y, dn = 1.0 / (dr * dr + di * di);
                const cplx sc = make_double2(dr * dn, -di * dn);           // 1 / (alpha - beta)
                for (int t = tid; t < ns; t += THREADS) vcur[t] = (t == 0) ? make_double2(1.0, 0.0) : cmul(pcur[t + 1], sc);
            }
            __syncthreads();
            // ---- (3) p = A0[S,S] v as column sums p[c] = sum_r conj(G[r][c]) v[r]  (A0 = the matrix at the start of the panel) -----
            {
                const int ncb = (ns + 31) >> 5;
                const int rsf = max(1, min(RSMAX, NWARPS / ncb));
                const int cb = warp % ncb, rs = warp / ncb;
                if (rs < rsf) {
                    const int c = j + 1 + cb * 32 + lane;
                    cplx acc = make_double2(0.0, 0.0);
                    if (c < n) {
                        const cplx* col = Gb + c;
                        int r = j + 1 + rs;
                        for (; r + 7 * rsf < n; r += 8 * rsf) {
                            cplx g[8];
#pragma unroll
                            for (int u = 0; u < 8; ++u) g[u] = col[(long long)(r + u * rsf) * n];
#pragma unroll
                            for (int u = 0; u < 8; ++u) acc = cfma(cconj(g[u]), vcur[r + u * rsf - (j + 1)], acc);
                        }
                        for (; r < n; r += rsf) acc = cfma(cconj(col[(long long)r * n]), vcur[r - (j + 1)], acc);
                        part[(size_t)rs * nmax + (c - (j + 1))] = acc;
                    }
                }
                // ---- (4) tw = W[S,:i]^H v, tv = V[S,:i]^H v: one warp per dot product (2 i of them) ------------------------------
                for (int job = warp; job < 2 * i; job += NWARPS) {
                    const int q = job >> 1;
                    const cplx* P = (job & 1) ? Vs : Ws;
                    double ar = 0.0, ai = 0.0;
                    for (int t = lane; t < ns; t += 32) {
                        const cplx x = P[(size_t)(j + 1 + t) * PL + q], y = vcur[t];
                        ar += x.x * y.x + x.y * y.y;            // conj(x) * y
                        ai += x.x * y.y - x.y * y.x;
                    }
                    ar = warp_sum(ar); ai = warp_sum(ai);
                    if (lane == 0) { if (job & 1) tv[q] = make_double2(ar, ai); else tw[q] = make_double2(ar, ai); }
                }
                __syncthreads();
                // ---- (5) p -= V tw + W tv ; w = tau p ; alpha2 = -1/2 tau (w^H v) ; w += alpha2 v --------------------------------
                double dot[2] = {0.0, 0.0};
                for (int t = tid; t < ns; t += THREADS) {
                    cplx p = part[t];
                    for (int g = 1; g < rsf; ++g) p = cadd(p, part[(size_t)g * nmax + t]);
                    const cplx* vr = Vs + (size_t)(j + 1 + t) * PL;
                    const cplx* wr = Ws + (size_t)(j + 1 + t) * PL;
                    for (int q = 0; q < i; ++q) {
                        const cplx x = cmul(vr[q], tw[q]), y = cmul(wr[q], tv[q]);
                        p.x -= x.x + y.x; p.y -= x.y + y.y;
                    }
                    p = cmul(tau, p);
                    pcur[t] = p;
                    const cplx vv = vcur[t];
                    dot[0] += p.x * vv.x + p.y * vv.y;          // conj(p) * v
                    dot[1] += p.x * vv.y - p.y * vv.x;
                }
                block_sum<2>(dot, red);
                const cplx al = cmul(make_double2(-0.5 * tau.x, -0.5 * tau.y), make_double2(dot[0], dot[1]));
                // ---- (6) panel columns i of V and W; row j of G <- v_j (zero-padded on the left) ----------------------------------
                for (int t = tid; t < ns; t += THREADS) {
                    const cplx vv = vcur[t];
                    Vs[(size_t)(j + 1 + t) * PL + i] = vv;
                    Ws[(size_t)(j + 1 + t) * PL + i] = cfma(al, vv, pcur[t]);
                }
                cplx* rowj = Gb + (long long)j * n;
                for (int t = tid; t < n; t += THREADS) rowj[t] = (t <= j) ? make_double2(0.0, 0.0) : vcur[t - (j + 1)];
                if (tid == 0) { d[j] = dj; e[j] = beta; tauv[j] = tau; }
                __syncthreads();
            }
        }
        // ---- rank-2 nb update of the trailing block [T0, n): lower triangle in 32 x 32 tiles, mirrored --------------------------
        const int T0 = j0 + nb;
        const int nt = (n - T0 + 31) >> 5;
        const int ntiles = nt * (nt + 1) / 2;
        for (int tile = warp; tile < ntiles; tile += NWARPS) {
            int R = (int)((sqrt(8.0 * tile + 1.0) - 1.0) * 0.5);
            while ((R + 1) * (R + 2) / 2 <= tile) ++R;
            while (R * (R + 1) / 2 > tile) --R;
            const int C = tile - R * (R + 1) / 2;
            const int c = T0 + 32 * C + lane;
            const bool cok = c < n;
            const cplx* wc_p = Ws + (size_t)(cok ? c : T0) * PL;
            const cplx* vc_p = Vs + (size_t)(cok ? c : T0) * PL;
            for (int r0 = T0 + 32 * R; r0 < min(n, T0 + 32 * R + 32); r0 += 4) {
                cplx acc[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) acc[u] = make_double2(0.0, 0.0);
                for (int q = 0; q < nb; ++q) {
                    const cplx wc = cconj(wc_p[q]), vc = cconj(vc_p[q]);
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int r = min(r0 + u, n - 1);
                        acc[u] = cfma(Vs[(size_t)r * PL + q], wc, acc[u]);
                        acc[u] = cfma(Ws[(size_t)r * PL + q], vc, acc[u]);
                    }
                }
                if (cok) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int r = r0 + u;
                        if (r >= n) continue;
                        if (R == C && c > r) continue;             // diagonal tile: its upper half is written as the mirror of the lower
                        cplx g = Gb[(long long)r * n + c];
                        g.x -= acc[u].x; g.y -= acc[u].y;
                        if (r == c) g.y = 0.0;
                        Gb[(long long)r * n + c] = g;
                        if (r != c) Gb[(long long)c * n + r] = cconj(g);
                    }
                }
            }
        }
        __syncthreads();
    }
    if (tid == 0) { d[n - 1] = Gb[(long long)(n - 1) * n + (n - 1)].x; e[n - 1] = 0.0; tauv[n - 1] = make_double2(0.0, 0.0); }
    __syncthreads();
    for (int t = tid; t < n; t += THREADS) Gb[(long long)(n - 1) * n + t] = make_double2(0.0, 0.0);
}

// E^H = Z^T Qh^H, row by row: M[c][:] <- M[c][:] H_{n-2}^H ... H_0^H with H_j^H = I - conj(tau_j) v_j v_j^H, v_j = row j of G.
// PB reflectors at a time are staged in shared memory; a warp owns the rows c = warp, warp + NWARPS, ... and keeps the row it is
// working on in registers (element e = lane + 32 t), so a reflector costs one dot product (with a warp reduction) and one axpy.
// NE = ceil(nmax / 32) elements per lane.
template <int NE>
__global__ void __launch_bounds__(THREADS, 1) k_ormtr_rows(const cplx* __restrict__ G, long long sG, const cplx* __restrict__ tauin, long long s_vec,
                                                           cplx* M, long long sM, const int* ek, int esk, int pb, const int* mask,
                                                           const double* __restrict__ srcR = nullptr, long long sR = 0, const int* __restrict__ order = nullptr,
                                                           long long s_ord = 0, int top = 0) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int b = blockIdx.x;
    if (mask && !mask[b]) return;
    const int n = ek[(long long)b * esk];
    if (n <= 0 || (n == 1 && !srcR)) return;               // (a 1 x 1 matrix has no reflector; in "top" mode its row is still gathered)
    cplx* Vp = reinterpret_cast<cplx*>(smem_raw);          // [pb][NE * 32]
    cplx* taus = Vp + (size_t)pb * NE * 32;                // [pb] conj(tau)
    const cplx* Gb = G + (long long)b * sG;
    const cplx* tauv = tauin + (long long)b * s_vec;
    cplx* Mb = M + (long long)b * sM;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int LD = NE * 32;
    // "top" mode (a split capped at max_bond < n can only keep the eigenvectors of the `top` largest eigenvalues): the rows order[0 ..
    // nrows) of the REAL eigenvector matrix of the tridiagonal (srcR, row c = eigenvector c) are gathered into rows 0 .. nrows of M and
    // only those are back-transformed -- half the reflector applications at chi = 128 of 2 chi = 256
    const int nrows = (srcR && top > 0) ? min(n, top) : n;
    if (srcR) {
        const double* Rb = srcR + (long long)b * sR;
        const int* ord = order + (long long)b * s_ord;
        for (int x = tid; x < nrows * n; x += THREADS) {
            const int c = x / n, e = x - c * n;
            Mb[x] = make_double2(Rb[(long long)ord[c] * n + e], 0.0);
        }
        __syncthreads();
    }
    if (n == 1) return;
    for (int jhi = n - 2; jhi >= 0; jhi -= pb) {
        const int jlo = max(0, jhi - pb + 1), cnt = jhi - jlo + 1;
        __syncthreads();                                    // previous panel fully consumed
        for (int x = tid; x < cnt * LD; x += THREADS) {
            const int q = x / LD, e = x - q * LD;
            Vp[x] = (e < n) ? Gb[(long long)(jlo + q) * n + e] : make_double2(0.0, 0.0);
        }
        if (tid < cnt) taus[tid] = cconj(tauv[jlo + tid]);
        __syncthreads();
        const int t0 = (jlo + 1) >> 5;                     // first 32-element chunk any reflector of the panel touches
        // two rows per warp and iteration: the reflector is read from shared memory once for both, and the two dot -> reduce -> axpy
        // chains are independent (the first version walked one row per warp and spent 40 % of its stalls on those shared-memory reads,
        // profiles/r02_ncu_heev_rows_v1.json)
        constexpr bool PAIR = NE <= 8;                      // 16 elements per lane (k > 256): two rows would not fit in 128 registers
        constexpr int RPW = PAIR ? 2 : 1;
        for (int c = RPW * warp; c < nrows; c += RPW * NWARPS) {
            const bool two = PAIR && (c + 1 < nrows);
            cplx ra[NE], rb[NE];
            cplx* mrow_a = Mb + (long long)c * n;
            cplx* mrow_b = mrow_a + n;
#pragma unroll
            for (int t = 0; t < NE; ++t) {
                const int e = lane + 32 * t;
                const bool ok = t >= t0 && e < n;
                ra[t] = ok ? mrow_a[e] : make_double2(0.0, 0.0);
                rb[t] = (ok && two) ? mrow_b[e] : make_double2(0.0, 0.0);
            }
            for (int q = cnt - 1; q >= 0; --q) {
                const cplx* v = Vp + (size_t)q * LD + lane;
                const int tq = (jlo + q + 1) >> 5;
                cplx vv[PAIR ? NE : 1];                      // the reflector stays in registers between dot and axpy when it fits
                // two partial sums per dot product (even / odd chunks): half the length of the dependent FMA chain in front of the
                // warp reduction (the chain, not the FP64 pipe, bounded this kernel: 42 % pipe utilisation in profiles/r02_ncu_target_v1.json)
                double sar[2] = {0.0, 0.0}, sai[2] = {0.0, 0.0}, sbr[2] = {0.0, 0.0}, sbi[2] = {0.0, 0.0};
#pragma unroll
                for (int t = 0; t < NE; ++t) {
                    if (t < tq) continue;
                    const cplx x = v[32 * t];
                    if constexpr (PAIR) vv[t] = x;
                    sar[t & 1] += ra[t].x * x.x - ra[t].y * x.y;
                    sai[t & 1] += ra[t].x * x.y + ra[t].y * x.x;
                    if constexpr (PAIR) {
                        sbr[t & 1] += rb[t].x * x.x - rb[t].y * x.y;
                        sbi[t & 1] += rb[t].x * x.y + rb[t].y * x.x;
                    }
                }
                sar[0] += sar[1]; sai[0] += sai[1]; sbr[0] += sbr[1]; sbi[0] += sbi[1];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    sar[0] += __shfl_xor_sync(0xffffffffu, sar[0], o); sai[0] += __shfl_xor_sync(0xffffffffu, sai[0], o);
                    if constexpr (PAIR) { sbr[0] += __shfl_xor_sync(0xffffffffu, sbr[0], o); sbi[0] += __shfl_xor_sync(0xffffffffu, sbi[0], o); }
                }
                const cplx fa = cmul(taus[q], make_double2(sar[0], sai[0])), fb = cmul(taus[q], make_double2(sbr[0], sbi[0]));   // conj(tau) (row . v)
#pragma unroll
                for (int t = 0; t < NE; ++t) {
                    if (t < tq) continue;
                    cplx x;
                    if constexpr (PAIR) x = vv[t]; else x = v[32 * t];
                    ra[t].x -= fa.x * x.x + fa.y * x.y;                        // f * conj(v)
                    ra[t].y -= fa.y * x.x - fa.x * x.y;
                    if constexpr (PAIR) {
                        rb[t].x -= fb.x * x.x + fb.y * x.y;
                        rb[t].y -= fb.y * x.x - fb.x * x.y;
                    }
                }
            }
#pragma unroll
            for (int t = 0; t < NE; ++t) {
                const int e = lane + 32 * t;
                if (t >= t0 && e < n) {
                    mrow_a[e] = ra[t];
                    if (two) mrow_b[e] = rb[t];
                }
            }
        }
    }
}

}  // namespace hetrd
