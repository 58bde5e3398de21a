// Batched complex128 contraction kernel for the TDVP hot path (sm_100a).
//
// Every contraction of the reference's tensordot chain (core/methods/tdvp/primitives.py:77-226,
// core/methods/decompositions.py:87-102) is a (grouped, K-segmented) complex GEMM over per-trajectory
// operands with ragged extents.  FP64 has no tcgen05 kind, so the tensor path is the warp-level
// DMMA (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4); one complex MAC = 4 DMMA issues with the sign of the
// Ai*Bi term folded into the A fragment (integer XOR on the sign bit, off the FP64 pipe).
//
// Tiling: CTA 64 x 128 complex, BK = 8, 4-stage cp.async (LDGSTS.128, one complex per copy, zero-fill
// for ragged edges), 8 warps as 2 x 4, warp tile 32 x 32 => 64 fp64 accumulators per thread.  Operands
// stay interleaved in shared memory so one LDS.128 feeds both the real and imaginary fragment; the row
// pitches (12 / 66 / 130 complex) make every quarter-warp LDS.128 hit 8 distinct 16-byte bank groups.
#include "yb_common.cuh"

namespace {

constexpr int BM = 64, BN = 128, BKT = 8, STAGES = 4;
constexpr int A_KIN_LD = BKT + 4;   // A stored [m][k]
constexpr int A_KOUT_LD = BM + 2;   // A stored [k][m]
constexpr int B_KIN_LD = BKT + 4;   // B stored [n][k]
constexpr int B_KOUT_LD = BN + 2;   // B stored [k][n]
constexpr int A_STAGE = (BM * A_KIN_LD > BKT * A_KOUT_LD) ? BM * A_KIN_LD : BKT * A_KOUT_LD;   // complex elements
constexpr int B_STAGE = (BN * B_KIN_LD > BKT * B_KOUT_LD) ? BN * B_KIN_LD : BKT * B_KOUT_LD;

enum { LAY_AK = 0, LAY_AM = 1 };   // A global layout: k contiguous / m contiguous
enum { LAY_BN = 0, LAY_BK = 1 };   // B global layout: n contiguous / k contiguous

struct Desc {
    const cplx* A; const cplx* B; cplx* C;
    int M, N, K, nseg;
    long long a_rs, a_cs, a_ss, b_rs, b_cs, b_ss, c_rs;
    bool conjB;
};

__device__ __forceinline__ bool make_desc(const YbGemmParams& P, int b, int g, Desc& d) {
    const int x0 = P.e0 ? P.e0[(long long)b * P.es0] : 0;
    const int x1 = P.e1 ? P.e1[(long long)b * P.es1] : 0;
    const int x2 = P.e2 ? P.e2[(long long)b * P.es2] : 0;
    const cplx* A = P.A + (long long)b * P.sA;
    const cplx* B = P.Bm + (long long)b * P.sB;
    cplx* C = P.C + (long long)b * P.sC;
    d.nseg = 1; d.a_ss = 0; d.b_ss = 0; d.conjB = false; d.a_cs = 1; d.b_cs = 1;
    switch (P.op) {
    case YB_OP_RCONTRACT: {           // x0 = l, x1 = r
        d.M = P.dphys * x0; d.K = x1; d.N = P.Dr * x1;
        d.A = A; d.a_rs = x1;
        d.B = B; d.b_rs = (long long)P.Dr * x1;
        d.C = C; d.c_rs = (long long)P.Dr * x1;
    } break;
    case YB_OP_LCONTRACT: {           // x0 = l, x1 = r ; group o
        d.M = x0; d.K = x0 * P.Dl; d.N = x1;
        d.A = A; d.a_rs = 1; d.a_cs = x0;
        d.B = B + (long long)g * x0 * P.Dl * x1; d.b_rs = x1;
        d.C = C + (long long)g * x0 * x1; d.c_rs = x1;
    } break;
    case YB_OP_ENVR3: {               // x0 = l, x1 = r ; segments o
        d.M = x0 * P.Dl; d.N = x0; d.K = x1; d.nseg = P.dphys;
        d.A = A; d.a_rs = x1; d.a_ss = (long long)x0 * P.Dl * x1;
        d.B = B; d.b_rs = 1; d.b_cs = x1; d.b_ss = (long long)x0 * x1; d.conjB = true;
        d.C = C; d.c_rs = x0;
    } break;
    case YB_OP_ENVL1: {               // x0 = l, x1 = r ; group o
        d.M = x0 * P.Dl; d.K = x0; d.N = x1;
        d.A = A; d.a_rs = x0;
        d.B = B + (long long)g * x0 * x1; d.b_rs = x1; d.conjB = true;
        d.C = C + (long long)g * x0 * P.Dl * x1; d.c_rs = x1;
    } break;
    case YB_OP_ENVL3: {               // x0 = l, x1 = r
        d.M = x1; d.K = P.dphys * x0; d.N = P.Dr * x1;
        d.A = A; d.a_rs = 1; d.a_cs = x1;
        d.B = B; d.b_rs = (long long)P.Dr * x1;
        d.C = C; d.c_rs = (long long)P.Dr * x1;
    } break;
    case YB_OP_MERGE: {               // x0 = l, x1 = c, x2 = r ; group s*d + s'
        const int s = g / P.dphys, sp = g % P.dphys;
        d.M = x0; d.K = x1; d.N = x2;
        d.A = A + (long long)s * x0 * x1; d.a_rs = x1;
        d.B = B + (long long)sp * x1 * x2; d.b_rs = x2;
        d.C = C + (long long)g * x0 * x2; d.c_rs = x2;
    } break;
    case YB_OP_ABSORB_R: {            // x0 = k_new, x1 = j_old, x2 = r ; group s
        d.M = x0; d.K = x1; d.N = x2;
        d.A = A; d.a_rs = x1;
        d.B = B + (long long)g * x1 * x2; d.b_rs = x2;
        d.C = C + (long long)g * x0 * x2; d.c_rs = x2;
    } break;
    case YB_OP_ABSORB_L: {            // x0 = l, x1 = j_old, x2 = k_new ; group s
        d.M = x0; d.K = x1; d.N = x2;
        d.A = A + (long long)g * x0 * x1; d.a_rs = x1;
        d.B = B; d.b_rs = x2;
        d.C = C + (long long)g * x0 * x2; d.c_rs = x2;
    } break;
    case YB_OP_GRAM: {                // x0 = k, x1 = len
        d.M = x0; d.N = x0; d.K = x1;
        d.A = A; d.a_rs = x1;
        d.B = B; d.b_rs = 1; d.b_cs = x1; d.conjB = true;
        d.C = C; d.c_rs = x0;
    } break;
    case YB_OP_EVEC: {                // x0 = k
        d.M = x0; d.N = x0; d.K = x0;
        d.A = A; d.a_rs = x0;
        d.B = B; d.b_rs = 1; d.b_cs = x0; d.conjB = true;
        d.C = C; d.c_rs = x0;
    } break;
    case YB_OP_ROTATE: {              // x0 = k, x1 = len
        d.M = x0; d.K = x0; d.N = x1;
        d.A = A; d.a_rs = x0;
        d.B = B; d.b_rs = x1;
        d.C = C; d.c_rs = x1;
    } break;
    case YB_OP_ROTATE_TOP: {          // x0 = k, x1 = len, x2 = rows of the compact eigenvector block
        d.M = min(x2, x0); d.K = x0; d.N = x1;
        d.A = A; d.a_rs = x0;
        d.B = B; d.b_rs = x1;
        d.C = C; d.c_rs = x1;
    } break;
    default: return false;
    }
    if (P.conj_extra) d.conjB = !d.conjB;
    return d.M > 0 && d.N > 0;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, bool valid) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" :: "r"(dst), "l"(gmem_src), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" :: "n"(N)); }

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ double flip_sign(double x) {
    return __hiloint2double(__double2hiint(x) ^ 0x80000000, __double2loint(x));
}

// One output element per thread and K walked with FP64 FMAs: the small-product kernel, and the path a tile CTA takes when the ACTUAL
// product is tiny although its slot is not (bonds of 11 in slots of 64: a 64 x 128 DMMA tile per trajectory would be ~99 % padding).
__device__ __forceinline__ void gemm_elements(const Desc& d, int first, int stride) {
    const int total = d.M * d.N;
    for (int idx = first; idx < total; idx += stride) {
        const int m = idx / d.N, n = idx - m * d.N;
        double cr = 0.0, ci = 0.0;
        for (int seg = 0; seg < d.nseg; ++seg) {
            const cplx* a = d.A + seg * d.a_ss + (long long)m * d.a_rs;
            const cplx* bp = d.B + seg * d.b_ss + (long long)n * d.b_cs;
            for (int k = 0; k < d.K; ++k) {
                const cplx x = a[(long long)k * d.a_cs];
                cplx y = bp[(long long)k * d.b_rs];
                if (d.conjB) y.y = -y.y;
                cr = fma(x.x, y.x, cr); cr = fma(-x.y, y.y, cr);
                ci = fma(x.x, y.y, ci); ci = fma(x.y, y.x, ci);
            }
        }
        d.C[(long long)m * d.c_rs + n] = make_double2(cr, ci);
    }
}
constexpr int TINY_MN = 1024;      // actual M * N up to which a tile CTA walks the product element-wise (one CTA per trajectory and group)

// RAGGED (A/B option "gemm_ragged", default off): the 8 x 8 sub-tiles of a warp are interleaved over the CTA tile (rows i*16 + wm*8,
// columns j*32 + wn*8 instead of wm*32 + i*8 / wn*32 + j*8) and sub-tiles that lie entirely beyond M or N issue no fragment loads and no
// DMMA.  With ragged bonds (chi ~ 113 of 128: M = 4*113 = 452, N = 3*113 = 339) the last tile row / column of a 64 x 128 CTA tile is
// mostly padding; interleaving spreads the surviving sub-tiles evenly over the 8 warps (2 per SM sub-partition), so the DMMA pipe of
// every sub-partition sees the same reduction.  Same sums in the same order per output element: results are bit-identical.
// WM = warp rows of the CTA: 2 -> 8 warps with 32 x 32 warp tiles (round 1: 230 registers, two warps per scheduler; ncu put ~45 % of the
// stall samples of the DMMA stream on the fixed issue latency between tensor instructions of the SAME warp, profiles/
// r02_ncu_gemm_rw_source_hotspots_v1.txt), 4 -> 16 warps with 16 x 32 warp tiles (<= 128 registers, four warps per scheduler to fill
// those slots).  Option "gemm_warps16" / YB_GEMM_WARPS16 selects; sums and their order per output element are the same.
template <int ALAY, int BLAY, bool RAGGED, int WM>
__global__ void __launch_bounds__(128 * WM, 1) k_gemm(const YbGemmParams P) {
    constexpr int NTHREADS = 128 * WM, MT = 8 / WM;          // m-tiles of 8 rows per warp
    const int b = blockIdx.z, g = blockIdx.y;
    if (P.mask && !P.mask[b]) return;
    Desc d;
    if (!make_desc(P, b, g, d)) return;
    if (d.M * d.N <= TINY_MN && !P.no_small) {               // tiny actual product in a large slot: no tile, no pipeline
        if (blockIdx.x != 0) return;
        if (P.flops && threadIdx.x == 0) atomicAdd(P.flops, 8ull * (unsigned long long)(d.M * d.N) * (unsigned long long)d.K * d.nseg);
        gemm_elements(d, threadIdx.x, NTHREADS);
        return;
    }
    const int tiles_n = (d.N + BN - 1) / BN;
    const int tiles_m = (d.M + BM - 1) / BM;
    if ((int)blockIdx.x >= tiles_m * tiles_n) return;
    const int m0 = ((int)blockIdx.x / tiles_n) * BM, n0 = ((int)blockIdx.x % tiles_n) * BN;
    if (P.flops && threadIdx.x == 0)
        atomicAdd(P.flops, 8ull * (unsigned long long)min(BM, d.M - m0) * (unsigned long long)min(BN, d.N - n0) * (unsigned long long)d.K * d.nseg);

    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx* As = reinterpret_cast<cplx*>(smem_raw);
    cplx* Bs = As + STAGES * A_STAGE;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3;
    const int kt_per_seg = (d.K + BKT - 1) / BKT;
    const int ktiles = kt_per_seg * d.nseg;

    auto load_stage = [&](int kt, int stage) {
        if (kt < ktiles) {
            const int seg = kt / kt_per_seg, k0 = (kt % kt_per_seg) * BKT;
            const cplx* Ab = d.A + seg * d.a_ss;
            const cplx* Bb = d.B + seg * d.b_ss;
            cplx* as = As + stage * A_STAGE;
            cplx* bs = Bs + stage * B_STAGE;
#pragma unroll
            for (int q = 0; q < (BM * BKT) / NTHREADS; ++q) {
                const int e = tid + q * NTHREADS;
                int m, k;
                if (ALAY == LAY_AK) { k = e % BKT; m = e / BKT; } else { m = e % BM; k = e / BM; }
                const bool ok = (m0 + m < d.M) && (k0 + k < d.K);
                const cplx* src = ok ? (Ab + (long long)(m0 + m) * d.a_rs + (long long)(k0 + k) * d.a_cs) : Ab;
                cplx* dst = (ALAY == LAY_AK) ? (as + m * A_KIN_LD + k) : (as + k * A_KOUT_LD + m);
                cp_async16(dst, src, ok);
            }
#pragma unroll
            for (int q = 0; q < (BN * BKT) / NTHREADS; ++q) {
                const int e = tid + q * NTHREADS;
                int n, k;
                if (BLAY == LAY_BN) { n = e % BN; k = e / BN; } else { k = e % BKT; n = e / BKT; }
                const bool ok = (n0 + n < d.N) && (k0 + k < d.K);
                const cplx* src = ok ? (Bb + (long long)(k0 + k) * d.b_rs + (long long)(n0 + n) * d.b_cs) : Bb;
                cplx* dst = (BLAY == LAY_BN) ? (bs + k * B_KOUT_LD + n) : (bs + n * B_KIN_LD + k);
                cp_async16(dst, src, ok);
            }
        }
        cp_async_commit();
    };

    double cr[MT][4][2], ci[MT][4][2];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { cr[i][j][0] = cr[i][j][1] = ci[i][j][0] = ci[i][j][1] = 0.0; }

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) load_stage(s, s);

    const int fr = lane >> 2, fk = lane & 3;
    int rbase[MT], cbase[4];
    bool ron[MT], con[4];
#pragma unroll
    for (int i = 0; i < MT; ++i) {
        rbase[i] = i * (8 * WM) + wm * 8;
        ron[i] = !RAGGED || (m0 + rbase[i] < d.M);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        cbase[i] = i * 32 + wn * 8;
        con[i] = !RAGGED || (n0 + cbase[i] < d.N);
    }
    for (int kt = 0; kt < ktiles; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        load_stage(kt + STAGES - 1, (kt + STAGES - 1) % STAGES);
        const cplx* as = As + (kt % STAGES) * A_STAGE;
        const cplx* bs = Bs + (kt % STAGES) * B_STAGE;
#pragma unroll
        for (int kk = 0; kk < BKT; kk += 4) {
            double ar[MT], ai[MT], an[MT], br[4], bi[4];
#pragma unroll
            for (int i = 0; i < MT; ++i) {
                if (RAGGED && !ron[i]) { ar[i] = ai[i] = an[i] = 0.0; continue; }
                const int row = (RAGGED ? rbase[i] : wm * (8 * MT) + i * 8) + fr, k = kk + fk;
                const cplx v = (ALAY == LAY_AK) ? as[row * A_KIN_LD + k] : as[k * A_KOUT_LD + row];
                ar[i] = v.x; ai[i] = v.y; an[i] = flip_sign(v.y);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (RAGGED && !con[j]) { br[j] = bi[j] = 0.0; continue; }
                const int n = (RAGGED ? cbase[j] : wn * 32 + j * 8) + fr, k = kk + fk;
                const cplx v = (BLAY == LAY_BN) ? bs[k * B_KOUT_LD + n] : bs[n * B_KIN_LD + k];
                br[j] = v.x; bi[j] = d.conjB ? flip_sign(v.y) : v.y;
            }
            // two passes: the second DMMA on an accumulator is issued a whole tile sweep after the first (the asm statements are
            // volatile, so source order is issue order; back to back, the dependent pair stalled on the tensor pipe's latency --
            // "wait" was 44 % of the stall samples on this line, profiles/r02_ncu_gemm_rw_source_hotspots_v1.txt)
#pragma unroll
            for (int i = 0; i < MT; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (RAGGED && !(ron[i] && con[j])) continue;          // warp-uniform: the whole 8 x 8 sub-tile is padding
                    dmma(cr[i][j][0], cr[i][j][1], ar[i], br[j]);
                    dmma(ci[i][j][0], ci[i][j][1], ar[i], bi[j]);
                }
#pragma unroll
            for (int i = 0; i < MT; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (RAGGED && !(ron[i] && con[j])) continue;
                    dmma(cr[i][j][0], cr[i][j][1], an[i], bi[j]);
                    dmma(ci[i][j][0], ci[i][j][1], ai[i], br[j]);
                }
        }
    }
    cp_async_wait<0>();

#pragma unroll
    for (int i = 0; i < MT; ++i) {
        const int row = m0 + (RAGGED ? rbase[i] : wm * (8 * MT) + i * 8) + fr;
        if (row >= d.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int col = n0 + (RAGGED ? cbase[j] : wn * 32 + j * 8) + 2 * fk;
            cplx* dst = d.C + (long long)row * d.c_rs + col;
            if (col < d.N) dst[0] = make_double2(cr[i][j][0], ci[i][j][0]);
            if (col + 1 < d.N) dst[1] = make_double2(cr[i][j][1], ci[i][j][1]);
        }
    }
}

// Small products (every M and N of the launch <= 64: chi <= 16 chains -- BASELINE configs 1 and 4 -- and the chain-end sites of the
// large ones): the 64 x 128 DMMA tile would be mostly padding and its pipeline prologue / epilogue a fixed ~5 us per CTA, i.e. ~300 us
// per launch for 2048 trajectories (profiles/r02_bench_C1_v1.json: 0.75 s per step at L = 10, chi = 16).  Here a thread owns one output
// element and walks K with FP64 FMAs; a warp's loads of A are broadcasts (same row), of B coalesced (consecutive columns), both served
// by L1 (the operands of a trajectory are a few KB).  Same descriptors, same sums per element in the same k order.
template <int DUMMY>
__global__ void __launch_bounds__(128) k_gemm_small(const YbGemmParams P) {
    const int b = blockIdx.z, g = blockIdx.y;
    if (P.mask && !P.mask[b]) return;
    Desc d;
    if (!make_desc(P, b, g, d)) return;
    const int total = d.M * d.N;
    if (P.flops && threadIdx.x == 0 && blockIdx.x == 0) atomicAdd(P.flops, 8ull * (unsigned long long)total * (unsigned long long)d.K * d.nseg);
    // grid-stride over the output elements: the grid is sized from the batch (yb_launch_gemm), not from the capacity of the product
    gemm_elements(d, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}

constexpr size_t GEMM_SMEM = (size_t)STAGES * (A_STAGE + B_STAGE) * sizeof(cplx);

template <int ALAY, int BLAY, bool RAGGED, int WM>
cudaError_t launch_r(const YbGemmParams& P, int maxM, int maxN, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(k_gemm<ALAY, BLAY, RAGGED, WM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    const int tiles = ((maxM + BM - 1) / BM) * ((maxN + BN - 1) / BN);
    dim3 grid(tiles, P.ngroups, P.nbatch);
    k_gemm<ALAY, BLAY, RAGGED, WM><<<grid, 128 * WM, GEMM_SMEM, st>>>(P);
    return cudaGetLastError();
}

template <int ALAY, int BLAY>
cudaError_t launch_t(const YbGemmParams& P, int maxM, int maxN, cudaStream_t st) {
    if (P.warps16) return P.ragged ? launch_r<ALAY, BLAY, true, 4>(P, maxM, maxN, st) : launch_r<ALAY, BLAY, false, 4>(P, maxM, maxN, st);
    return P.ragged ? launch_r<ALAY, BLAY, true, 2>(P, maxM, maxN, st) : launch_r<ALAY, BLAY, false, 2>(P, maxM, maxN, st);
}

#include "yb_gemm_rw.cuh"

}  // namespace

// RCONTRACT with the MPO tensor applied in the epilogue (yb_gemm_rw.cuh): 64 x 96 CTA tiles for right MPO bonds 3 and 6, else 64 x 128.
cudaError_t yb_launch_gemm_rw(int nbatch, const int* mask, const int* el, int esl, const int* er, int esr, int dphys, int Dl, int Dr, const cplx* X,
                              long long sX, const cplx* R, long long sR, const cplx* W, cplx* Out, long long sOut, int max_l, int max_r,
                              unsigned long long* flops, int bk16, int warps16, cudaStream_t st) {
    if (nbatch <= 0 || max_l <= 0 || max_r <= 0) return cudaSuccess;
    rw::Params P;
    P.nbatch = nbatch; P.mask = mask; P.el = el; P.er = er; P.esl = esl; P.esr = esr; P.dphys = dphys; P.Dl = Dl; P.Dr = Dr;
    P.X = X; P.R = R; P.W = W; P.Out = Out; P.sX = sX; P.sR = sR; P.sOut = sOut; P.flops = flops;
    (void)bk16;                                           // 16-deep k tiles measured no faster (profiles/r02_matvec_microbench_v1.json)
    if (warps16) return (Dr == 3 || Dr == 6) ? rw::launch<8, 4, 3, 4>(P, max_l, max_r, st) : rw::launch<8, 4, 4, 4>(P, max_l, max_r, st);
    return (Dr == 3 || Dr == 6) ? rw::launch<8, 4, 3, 2>(P, max_l, max_r, st) : rw::launch<8, 4, 4, 2>(P, max_l, max_r, st);
}

// maxM / maxN: upper bounds of the per-trajectory M and N (grid is sized for them; empty tiles exit).
cudaError_t yb_launch_gemm(const YbGemmParams& P, int maxM, int maxN, cudaStream_t st) {
    if (P.nbatch <= 0 || maxM <= 0 || maxN <= 0) return cudaSuccess;
    if (maxM <= P.small_max && maxN <= P.small_max && !P.no_small) {
        // one thread per capacity element.  Slots up to small_max x small_max only (default 32: chi <= 16 chains and chain ends): a 64 x 64
        // slot is either full (config 2: the DMMA tile wins) or nearly empty (config 4: a tile CTA walks a tiny actual product element-wise,
        // one CTA per trajectory and group instead of 32 mostly idle ones here)
        dim3 grid((maxM * maxN + 127) / 128, P.ngroups, P.nbatch);
        k_gemm_small<0><<<grid, 128, 0, st>>>(P);
        return cudaGetLastError();
    }
    switch (P.op) {
    case YB_OP_LCONTRACT:
    case YB_OP_ENVL3: return launch_t<LAY_AM, LAY_BN>(P, maxM, maxN, st);
    case YB_OP_ENVR3:
    case YB_OP_GRAM:
    case YB_OP_EVEC: return launch_t<LAY_AK, LAY_BK>(P, maxM, maxN, st);
    default: return launch_t<LAY_AK, LAY_BN>(P, maxM, maxN, st);
    }
}
