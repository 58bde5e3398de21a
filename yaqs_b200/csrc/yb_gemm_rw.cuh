// project_site, first two steps in ONE kernel (primitives.py:201-202 / :105-106):
//     T1[(u,l),(v,r')] = sum_r X[(u,l),r] R[r,(v,r')]          the GEMM of YB_OP_RCONTRACT (DMMA main loop of yb_gemm.cu)
//     T2[x,l,y,r']     = sum_(u,v) W[x,u,y,v] T1[u,l,v,r']      the MPO-tensor step, applied to the tile while it is on chip
// Round 1 ran them as k_gemm + k_wapply: a d~ D chi^2 intermediate written and re-read per matvec (re-read ~3.6x because k_wapply
// gathers one (x, y) slab per block) -- 7 % of the C3 step for an element-wise combination of values the GEMM already holds.
//
// The W step mixes ALL physical indices u and ALL right MPO bonds v at a fixed (l, r'), so the CTA tile is laid out in a virtual
// index space: tile row m = (u, li) with li < TL = 64 / d~, tile column n = (v, ri) with ri < TR = the largest multiple of 8 with
// Dr * TR <= 128.  Rows / columns are gathered by the cp.async loads (address arithmetic only), the main loop is the 64 x 128 DMMA
// loop of k_gemm, and the epilogue goes through shared memory (the pipeline buffers are free by then): the accumulator tile is
// parked as T[64][128] and the MPO step runs as a second small GEMM on the tensor pipe (see the epilogue).  A first version did
// the combination with scalar FMAs per output element and was SLOWER than the two-kernel chain it replaces (1.94 ms against 0.93 +
// 0.44 ms at B = 128, chi = 128, profiles/r02_matvec_microbench_v1.json): with one CTA per SM nothing overlaps the epilogue.
#pragma once

namespace rw {

constexpr int BM = 64;
constexpr int W_MAX = 4 * 4 * 8 * 8;             // MPO tensor entries held in shared memory (d~ <= 4, bonds <= 8)
constexpr int W2_MAX = 32 * 33;                  // the same tensor as a zero-padded (d~ Dl) x (d~ Dr) matrix, pitch Kp + 1

// NT = 8-column tiles per warp: CTA tile 64 x (32 NT).  NT = 3 (64 x 96) when the right MPO bond is 3 or 6 -- TR = 32 / 16 then fills
// the tile exactly and chi = 128 splits into whole column tiles (the 64 x 128 tile would run 4 column tiles of 40 with the last one
// 20 % full, or waste a quarter of its columns); NT = 4 otherwise.
template <int BKT, int STAGES, int NT>
struct Cfg {
    static constexpr int BN = 32 * NT;
    static constexpr int T_LD = BN + 2;                          // epilogue tile pitch (complex)
    static constexpr int A_LD = BKT + 4;                         // A stored [m][k]
    static constexpr int B_LD = BN + 2;                          // B stored [k][n]
    static constexpr int A_STAGE = BM * A_LD, B_STAGE = BKT * B_LD;
    static constexpr size_t PIPE = (size_t)STAGES * (A_STAGE + B_STAGE);
    static constexpr size_t TILE = (size_t)BM * T_LD;
    static constexpr size_t SMEM = ((PIPE > TILE ? PIPE : TILE) + W_MAX + W2_MAX) * sizeof(cplx) + (32 + 32) * sizeof(long long);
};

struct Params {
    int nbatch;
    const int* mask;
    const int* el; const int* er; int esl, esr;
    int dphys, Dl, Dr;                 // d~ (= nx = nu), left MPO bond (ny, kept), right MPO bond (nv, contracted)
    const cplx* X; const cplx* R; const cplx* W; cplx* Out;
    long long sX, sR, sOut;
    unsigned long long* flops;
};

template <int BKT, int STAGES, int NT, int WM>
__global__ void __launch_bounds__(128 * WM, 1) k_gemm_rw(const Params P) {
    using C = Cfg<BKT, STAGES, NT>;
    constexpr int BN = C::BN, T_LD = C::T_LD;
    constexpr int NTHREADS = 128 * WM, MT = 8 / WM;          // WM warp rows x 4 warp columns; warp tile (8 MT) x (8 NT), see k_gemm
    const int b = blockIdx.z;
    if (P.mask && !P.mask[b]) return;
    const int chil = P.el[(long long)b * P.esl], chir = P.er[(long long)b * P.esr];
    if (chil <= 0 || chir <= 0) return;
    const int dph = P.dphys, Dr = P.Dr, Dl = P.Dl;
    const int TL = BM / dph, TR = (BN / Dr) & ~7;
    const int tiles_r = (chir + TR - 1) / TR, tiles_l = (chil + TL - 1) / TL;
    if ((int)blockIdx.x >= tiles_l * tiles_r) return;
    const int l0 = ((int)blockIdx.x / tiles_r) * TL, r0 = ((int)blockIdx.x % tiles_r) * TR;
    const int nl = min(TL, chil - l0), nr = min(TR, chir - r0);
    if (P.flops && threadIdx.x == 0)
        atomicAdd(P.flops, 8ull * (unsigned long long)(dph * nl) * (unsigned long long)(Dr * nr) * (unsigned long long)chir +
                               8ull * (unsigned long long)(dph * dph * Dl * Dr) * (unsigned long long)nl * (unsigned long long)nr);

    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx* As = reinterpret_cast<cplx*>(smem_raw);
    cplx* Bs = As + STAGES * C::A_STAGE;
    cplx* Ws = reinterpret_cast<cplx*>(smem_raw) + (C::PIPE > C::TILE ? C::PIPE : C::TILE);
    const cplx* Xb = P.X + (long long)b * P.sX;
    const cplx* Rb = P.R + (long long)b * P.sR;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3;
    const int K = chir;
    const int ktiles = (K + BKT - 1) / BKT;
    const int nw = dph * dph * Dl * Dr;
    for (int e = tid; e < nw; e += NTHREADS) Ws[e] = P.W[e];

    // per-thread gather offsets of the loads (fixed over the k loop)
    constexpr int AQ = (BM * BKT + NTHREADS - 1) / NTHREADS, BQ = (BN * BKT + NTHREADS - 1) / NTHREADS;
    long long a_off[AQ]; int a_k[AQ], a_dst[AQ]; bool a_ok[AQ];
#pragma unroll
    for (int q = 0; q < AQ; ++q) {
        const int e = tid + q * NTHREADS;
        const int k = e % BKT, m = (e / BKT) % BM;
        const int u = m / TL, li = m - u * TL;
        a_ok[q] = li < nl && e < BM * BKT;
        a_off[q] = ((long long)u * chil + l0 + li) * chir;
        a_k[q] = k;
        a_dst[q] = m * C::A_LD + k;
    }
    long long b_off[BQ]; int b_k[BQ], b_dst[BQ]; bool b_ok[BQ];
#pragma unroll
    for (int q = 0; q < BQ; ++q) {
        const int e = tid + q * NTHREADS;
        const int n = e % BN, k = (e / BN) % BKT;
        const int v = n / TR, ri = n - v * TR;
        const bool in_tile = e < BN * BKT;
        b_ok[q] = v < Dr && ri < nr && in_tile;
        b_off[q] = (long long)v * chir + r0 + ri;
        b_k[q] = k;
        b_dst[q] = k * C::B_LD + n;
    }
    const long long r_rs = (long long)Dr * chir;

    auto load_stage = [&](int kt, int stage) {
        if (kt < ktiles) {
            const int k0 = kt * BKT;
            cplx* as = As + stage * C::A_STAGE;
            cplx* bs = Bs + stage * C::B_STAGE;
#pragma unroll
            for (int q = 0; q < AQ; ++q) {
                if (tid + q * NTHREADS >= BM * BKT) continue;           // (only when the tile does not divide among the threads)
                const bool ok = a_ok[q] && (k0 + a_k[q] < K);
                cp_async16(as + a_dst[q], ok ? (Xb + a_off[q] + k0 + a_k[q]) : Xb, ok);
            }
#pragma unroll
            for (int q = 0; q < BQ; ++q) {
                if (tid + q * NTHREADS >= BN * BKT) continue;
                const bool ok = b_ok[q] && (k0 + b_k[q] < K);
                cp_async16(bs + b_dst[q], ok ? (Rb + (long long)(k0 + b_k[q]) * r_rs + b_off[q]) : Rb, ok);
            }
        }
        cp_async_commit();
    };

    double cr[MT][NT][2], ci[MT][NT][2];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) { cr[i][j][0] = cr[i][j][1] = ci[i][j][0] = ci[i][j][1] = 0.0; }

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) load_stage(s, s);

    const int fr = lane >> 2, fk = lane & 3;
    for (int kt = 0; kt < ktiles; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        load_stage(kt + STAGES - 1, (kt + STAGES - 1) % STAGES);
        const cplx* as = As + (kt % STAGES) * C::A_STAGE;
        const cplx* bs = Bs + (kt % STAGES) * C::B_STAGE;
#pragma unroll
        for (int kk = 0; kk < BKT; kk += 4) {
            double ar[MT], ai[MT], an[MT], br[NT], bi[NT];
#pragma unroll
            for (int i = 0; i < MT; ++i) {
                const cplx v = as[(wm * (8 * MT) + i * 8 + fr) * C::A_LD + kk + fk];
                ar[i] = v.x; ai[i] = v.y; an[i] = flip_sign(v.y);
            }
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                const cplx v = bs[(kk + fk) * C::B_LD + wn * (8 * NT) + j * 8 + fr];
                br[j] = v.x; bi[j] = v.y;
            }
#pragma unroll
            for (int i = 0; i < MT; ++i)                      // two passes: dependent DMMAs a tile sweep apart (see k_gemm)
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    dmma(cr[i][j][0], cr[i][j][1], ar[i], br[j]);
                    dmma(ci[i][j][0], ci[i][j][1], ar[i], bi[j]);
                }
#pragma unroll
            for (int i = 0; i < MT; ++i)
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    dmma(cr[i][j][0], cr[i][j][1], an[i], bi[j]);
                    dmma(ci[i][j][0], ci[i][j][1], ai[i], br[j]);
                }
        }
    }
    cp_async_wait<0>();
    __syncthreads();                                  // every warp is done with the pipeline buffers

    // ---- epilogue: the MPO-tensor step as a second, small GEMM on the tensor pipe ----------------------------------------------
    //   T2[(x,y), pos] = sum_(u,v) W2[(x,y),(u,v)] T1[(u,v), pos],   pos = (li, ri), W2 = the MPO tensor read as a (d~ Dl) x (d~ Dr) matrix
    // The accumulator tile is parked in shared memory (the pipeline buffers are free), W2 sits zero-padded to multiples of 8 x 4 next
    // to it, and every warp takes 8-position column tiles (8 consecutive ri of one li): B fragments are gathered from the parked
    // tile, A fragments from W2, C fragments go straight to global memory (2 adjacent r' per lane: 32-byte stores).  Cost: d~ Dl x
    // d~ Dr x TL TR complex multiply-adds per CTA, ~10 % of the main loop at C3 sizes, instead of a second kernel and an HBM round trip.
    cplx* Ts = reinterpret_cast<cplx*>(smem_raw);
#pragma unroll
    for (int i = 0; i < MT; ++i) {
        const int row = wm * (8 * MT) + i * 8 + fr;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const int col = wn * (8 * NT) + j * 8 + 2 * fk;
            Ts[row * T_LD + col] = make_double2(cr[i][j][0], ci[i][j][0]);
            Ts[row * T_LD + col + 1] = make_double2(cr[i][j][1], ci[i][j][1]);
        }
    }
    const int Mw = dph * Dl, Kw = dph * Dr;                 // W2 extents
    const int Mp = (Mw + 7) & ~7, Kp = (Kw + 3) & ~3;
    const int W2_LD = Kp + 1;
    cplx* W2 = Ws + W_MAX;                                  // [Mp][Kp + 1], zero-padded
    int* koff = reinterpret_cast<int*>(W2 + W2_MAX);        // [Kp]  offset of (u, v) in the parked tile, -1 beyond Kw   (64 ints)
    long long* moff = reinterpret_cast<long long*>(koff + 64);   // [Mp]  offset of (x, y) in the output, -1 beyond Mw          (32)
    for (int e = tid; e < Mp * W2_LD; e += NTHREADS) {
        const int m = e / W2_LD, k = e - m * W2_LD;
        cplx w = make_double2(0.0, 0.0);
        if (m < Mw && k < Kw) {
            const int x = m / Dl, y = m - x * Dl, u = k / Dr, v = k - u * Dr;
            w = Ws[((x * dph + u) * Dl + y) * Dr + v];
        }
        W2[e] = w;
    }
    if (tid < Kp) { const int u = tid / Dr, v = tid - u * Dr; koff[tid] = tid < Kw ? u * TL * T_LD + v * TR : -1; }
    if (tid >= 64 && tid < 64 + Mp) {
        const int m = tid - 64, x = m / Dl, y = m - x * Dl;
        moff[m] = m < Mw ? ((long long)x * chil * Dl + y) * chir : -1;
    }
    __syncthreads();
    cplx* Ob = P.Out + (long long)b * P.sOut + (long long)l0 * Dl * chir + r0;
    const int tiles_per_row = TR >> 3;                      // 8-position tiles per li
    const int ntile = nl * tiles_per_row;                   // only the rows this tile owns
    for (int nt_ = warp; nt_ < ntile; nt_ += NTHREADS / 32) {
        const int li = nt_ / tiles_per_row, ri0 = (nt_ - li * tiles_per_row) << 3;
        if (ri0 >= nr) continue;                            // warp-uniform: nothing to store from this tile
        const cplx* tpos = Ts + li * T_LD + ri0 + fr;
        const int ri = ri0 + 2 * fk;
        cplx* orow = Ob + (long long)li * Dl * chir + ri;
        for (int m0 = 0; m0 < Mp; m0 += 8) {
            double c_r0 = 0.0, c_r1 = 0.0, c_i0 = 0.0, c_i1 = 0.0;
            for (int k0 = 0; k0 < Kp; k0 += 4) {
                const cplx av = W2[(m0 + fr) * W2_LD + k0 + fk];
                const int off = koff[k0 + fk];
                const cplx bv = off >= 0 ? tpos[off] : make_double2(0.0, 0.0);
                dmma(c_r0, c_r1, av.x, bv.x);
                dmma(c_i0, c_i1, av.x, bv.y);
                dmma(c_r0, c_r1, flip_sign(av.y), bv.y);
                dmma(c_i0, c_i1, av.y, bv.x);
            }
            const long long mo = moff[m0 + fr];
            if (mo >= 0 && ri < nr) {
                cplx* dst = orow + mo;
                dst[0] = make_double2(c_r0, c_i0);
                if (ri + 1 < nr) dst[1] = make_double2(c_r1, c_i1);
            }
        }
    }
}

template <int BKT, int STAGES, int NT, int WM>
cudaError_t launch(const Params& P, int max_l, int max_r, cudaStream_t st) {
    using C = Cfg<BKT, STAGES, NT>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(k_gemm_rw<BKT, STAGES, NT, WM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    const int TL = BM / P.dphys, TR = (C::BN / P.Dr) & ~7;
    const int tiles = ((max_l + TL - 1) / TL) * ((max_r + TR - 1) / TR);
    dim3 grid(tiles, 1, P.nbatch);
    k_gemm_rw<BKT, STAGES, NT, WM><<<grid, 128 * WM, C::SMEM, st>>>(P);
    return cudaGetLastError();
}

}  // namespace rw
