// extern "C" boundary of libyaqs_b200.so (declared in include/yaqs_b200.h).  Plain pointers and sizes only.
#include "../../include/yaqs_b200.h"
#include "yb_common.cuh"
#include <atomic>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

// unity build: the kernels live in the same translation unit (no -rdc needed)
#include "yb_gemm.cu"
#include "yb_ops.cu"
#include "yb_svd.cu"
#include "yb_heev.cu"

namespace {
thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
unsigned long long* g_flop_counter = nullptr;

int fail(const char* where, cudaError_t e) {
    snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
    return -1;
}
int fail_msg(const char* msg) {
    snprintf(g_err, sizeof(g_err), "%s", msg);
    return -2;
}
inline int done(const char* where) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : fail(where, e);
}
inline int grid_for(long long n, int threads, int cap = 4096) {
    long long g = (n + threads - 1) / threads;
    if (g < 1) g = 1;
    if (g > cap) g = cap;
    return (int)g;
}
// CTAs per trajectory for the element-wise kernels (all of them grid-stride over their elements).  The element count is a CAPACITY
// (the largest admissible bond dimensions), the actual extents are known on the device only: at 2048 trajectories with bonds of 11
// inside slots of 64 (config 4) a grid of 64 CTAs per trajectory is 131 k CTAs of which 129 k find nothing to do -- ~40 us per
// launch, 14 launches per gauge shift (profiles/r02_bench_C4_v2_graphs.json: 4.6 s per Trotter step at a quarter of the host calls).
// Enough CTAs to fill the GPU a few times over (8 per SM in total, 256 threads each: these kernels move a few bytes per element),
// never more than the capacity needs.  Measured on config 4 (`gpurun_out` calls 24-26): 8 per SM 2055, 16 per SM 1524 (with 64 x 64 slots on the tile kernel), 64 per SM 1635 trajectory-steps/s;
// compute-heavy element kernels must not be capped this way (config 2's 64 x 64 x 320 products on one CTA per trajectory: 149.5 -> 67.6),
// which is why the small-product GEMM keeps its full grid and 64 x 64 slots go to the tile kernel.
inline int grid_batch(long long n, int threads, int nbatch, int cap = 64) {
    const int need = grid_for(n, threads, cap);
    int want = (8 * 148 + nbatch - 1) / (nbatch > 0 ? nbatch : 1);
    if (want < 1) want = 1;
    return need < want ? need : want;
}
}  // namespace

extern "C" {

const char* yb_last_error(void) { return g_err; }
int yb_version(void) { return 100; }
int64_t yb_launch_count(int reset) {
    long long v = g_launches.load();
    if (reset) g_launches.store(0);
    return v;
}

// Device bytes a batch of `nbatch` trajectories needs (host arithmetic only; mirrors engine.BatchLayout / BatchedTJM.__init__).
int yb_workspace_bytes(int length, int nbatch, int dphys, const int32_t* cap, const int32_t* mpo_bond, int order, int two_envs, int gram,
                       yb_workspace* out) {
    if (!cap || !mpo_bond || !out || length < 1 || nbatch < 1 || dphys < 1) return fail_msg("yb_workspace_bytes: bad arguments");
    const int64_t C = 16, B = nbatch, d = dphys;
    int64_t state = 0, env = 0, pair = 1, single = 1, cmax = 1, dmax = 1;
    for (int i = 0; i < length; ++i) {
        state += d * cap[i] * (int64_t)cap[i + 1];
        single = single > (int64_t)cap[i] * cap[i + 1] ? single : (int64_t)cap[i] * cap[i + 1];
        if (i + 2 <= length) pair = pair > (int64_t)cap[i] * cap[i + 2] ? pair : (int64_t)cap[i] * cap[i + 2];
    }
    for (int j = 0; j <= length; ++j) {
        env += (int64_t)cap[j] * mpo_bond[j] * cap[j];
        cmax = cmax > cap[j] ? cmax : cap[j];
        dmax = dmax > mpo_bond[j] ? dmax : mpo_bond[j];
    }
    if (length < 2) pair = 1;
    const int64_t theta_slot = d * d * pair;
    const int64_t nloc = theta_slot > d * single ? theta_slot : d * single;
    const int64_t t_slot = nloc * dmax, sd = d * cmax, sq = sd * sd;
    memset(out, 0, sizeof(*out));
    out->state = C * state * B * (order == 2 ? 2 : 1) + 4 * B * (length + 1) * (order == 2 ? 2 : 1);
    out->environments = C * env * B * (two_envs ? 2 : 1);
    out->theta = C * B * theta_slot;
    out->matvec = C * B * 2 * t_slot;
    out->krylov = C * B * (26 + 1) * nloc + B * (2 * 32 * 8 + 32 * 16);
    out->svd = C * B * 4 * sq + 8 * B * sd + 4 * B * 8;
    const bool use_gram = gram && sd >= 33 && sd <= 512;
    out->heev = use_gram ? (C * B * 4 * sq + 8 * B * sq + 2 * 8 * B * sd + C * B * sd + 4 * B * sd) : 0;
    out->small = C * B * 4 * length + 64 * B;
    out->total = out->state + out->environments + out->theta + out->matvec + out->krylov + out->svd + out->heev + out->small;
    return 0;
}

void yb_set_flop_counter(void* device_u64) { g_flop_counter = (unsigned long long*)device_u64; }

// diagnostic switches (A/B runs and bisecting a failure); the defaults are the production route
static std::atomic<int> g_opt_heev_ql{getenv("YB_HEEV_QL") != nullptr ? 1 : 0};
static std::atomic<int> g_opt_heev_ring{getenv("YB_HEEV_RING") != nullptr ? atoi(getenv("YB_HEEV_RING")) : -1};

static std::atomic<int> g_opt_wapply_onepass{getenv("YB_WAPPLY_ONEPASS") != nullptr ? atoi(getenv("YB_WAPPLY_ONEPASS")) : 0};
static std::atomic<int> g_opt_gemm_bk16{getenv("YB_GEMM_BK16") != nullptr ? atoi(getenv("YB_GEMM_BK16")) : 0};
static std::atomic<int> g_opt_gemm_no_small{getenv("YB_GEMM_NO_SMALL") != nullptr ? atoi(getenv("YB_GEMM_NO_SMALL")) : 0};
static std::atomic<int> g_opt_gemm_small_max{getenv("YB_GEMM_SMALL_MAX") != nullptr ? atoi(getenv("YB_GEMM_SMALL_MAX")) : 32};
static std::atomic<int> g_opt_gemm_warps16{getenv("YB_GEMM_WARPS16") != nullptr ? atoi(getenv("YB_GEMM_WARPS16")) : 0};
static std::atomic<int> g_opt_gemm_ragged{getenv("YB_GEMM_RAGGED") != nullptr ? atoi(getenv("YB_GEMM_RAGGED")) : 0};

int yb_set_option(const char* name, int value) {
    if (!name) return fail_msg("yb_set_option: null name");
    if (!strcmp(name, "gemm_ragged")) { g_opt_gemm_ragged = value != 0; return 0; }
    if (!strcmp(name, "gemm_bk16")) { g_opt_gemm_bk16 = value != 0; return 0; }
    if (!strcmp(name, "gemm_warps16")) { g_opt_gemm_warps16 = value != 0; return 0; }
    if (!strcmp(name, "gemm_no_small")) { g_opt_gemm_no_small = value != 0; return 0; }
    if (!strcmp(name, "gemm_small_max")) { g_opt_gemm_small_max = value; return 0; }
    if (!strcmp(name, "wapply_onepass")) { g_opt_wapply_onepass = value != 0; return 0; }
    if (!strcmp(name, "heev_ql")) { g_opt_heev_ql = value != 0; return 0; }
    if (!strcmp(name, "heev_ring")) { g_opt_heev_ring = value; return 0; }
    return fail_msg("yb_set_option: unknown option");
}

int yb_gemm(int op, int nbatch, int ngroups, const int32_t* mask, yb_extent e0, yb_extent e1, yb_extent e2, int dphys, int d_left, int d_right,
            const void* a, int64_t stride_a, const void* b, int64_t stride_b, void* c, int64_t stride_c, int conj_b, int max_m, int max_n,
            void* stream) {
    if (op < 0 || op > YB_OP_LAST) return fail_msg("yb_gemm: unknown op");
    if (nbatch <= 0) return 0;
    YbGemmParams P;
    memset(&P, 0, sizeof(P));
    P.op = op; P.nbatch = nbatch; P.ngroups = ngroups; P.mask = mask;
    P.e0 = e0.ptr; P.es0 = e0.stride; P.e1 = e1.ptr; P.es1 = e1.stride; P.e2 = e2.ptr; P.es2 = e2.stride;
    P.dphys = dphys; P.Dl = d_left; P.Dr = d_right;
    P.A = (const cplx*)a; P.Bm = (const cplx*)b; P.C = (cplx*)c; P.sA = stride_a; P.sB = stride_b; P.sC = stride_c;
    P.conj_extra = conj_b;
    P.ragged = g_opt_gemm_ragged.load(std::memory_order_relaxed);
    P.warps16 = g_opt_gemm_warps16.load(std::memory_order_relaxed);
    P.no_small = g_opt_gemm_no_small.load(std::memory_order_relaxed);
    P.small_max = g_opt_gemm_small_max.load(std::memory_order_relaxed);
    P.flops = g_flop_counter;
    cudaError_t e = yb_launch_gemm(P, max_m, max_n, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail("yb_gemm", e);
    return done("yb_gemm");
}

int yb_rcontract_mpo(int nbatch, const int32_t* mask, yb_extent el, yb_extent er, int dphys, int d_left, int d_right, const void* x, int64_t stride_x,
                     const void* renv, int64_t stride_r, const void* wt, void* out, int64_t stride_out, int max_l, int max_r, void* stream) {
    if (nbatch <= 0) return 0;
    if (dphys != 1 && dphys != 2 && dphys != 4) return fail_msg("yb_rcontract_mpo: dphys must be 1, 2 or 4");
    if (d_left < 1 || d_right < 1 || d_left > 8 || d_right > 8)
        return fail_msg("yb_rcontract_mpo: MPO bond dimensions out of range (1 .. 8)");
    cudaError_t e = yb_launch_gemm_rw(nbatch, mask, el.ptr, el.stride, er.ptr, er.stride, dphys, d_left, d_right, (const cplx*)x, stride_x,
                                      (const cplx*)renv, stride_r, (const cplx*)wt, (cplx*)out, stride_out, max_l, max_r, g_flop_counter,
                                      g_opt_gemm_bk16.load(std::memory_order_relaxed), g_opt_gemm_warps16.load(std::memory_order_relaxed),
                                      (cudaStream_t)stream);
    if (e != cudaSuccess) return fail("yb_rcontract_mpo", e);
    return done("yb_rcontract_mpo");
}

int yb_wapply(int nbatch, const int32_t* mask, yb_extent el, yb_extent er, int nx, int nu, int ny, int nv, const void* wt, const void* in,
              int64_t stride_in, void* out, int64_t stride_out, int max_lr, void* stream) {
    if (nbatch <= 0 || max_lr <= 0) return 0;
    if (g_opt_wapply_onepass.load(std::memory_order_relaxed) && nx <= wapply::MAX_NX && nu <= wapply::MAX_NX && ny <= wapply::MAX_NY &&
        nv <= wapply::MAX_NY) {
        dim3 grid1((max_lr + 127) / 128, 1, nbatch);
        k_wapply_onepass<<<grid1, 128, 0, (cudaStream_t)stream>>>(nx, nu, ny, nv, (const cplx*)wt, (const cplx*)in, stride_in, (cplx*)out, stride_out,
                                                                  el.ptr, el.stride, er.ptr, er.stride, mask);
        return done("yb_wapply");
    }
    dim3 grid((max_lr + 255) / 256, nx * ny, nbatch);
    k_wapply<<<grid, 256, 0, (cudaStream_t)stream>>>(nx, nu, ny, nv, (const cplx*)wt, (const cplx*)in, stride_in, (cplx*)out, stride_out, el.ptr,
                                                      el.stride, er.ptr, er.stride, mask);
    return done("yb_wapply");
}

int yb_env_identity(int nbatch, const int32_t* mask, yb_extent echi, int dmpo, void* env, int64_t stride_env, int max_chi, void* stream) {
    if (nbatch <= 0) return 0;
    dim3 grid(grid_batch((long long)max_chi * dmpo * max_chi, 256, nbatch), nbatch);
    k_env_identity<<<grid, 256, 0, (cudaStream_t)stream>>>(dmpo, (cplx*)env, stride_env, echi.ptr, echi.stride, mask);
    return done("yb_env_identity");
}

int yb_lanczos_init(int nbatch, const int32_t* mask, yb_extent el, yb_extent er, int dphys, const void* x, int64_t stride_x, void* basis,
                    int64_t stride_basis, int64_t stride_vec, double* nrm, int32_t* active, int32_t* kdone, int32_t* n_zero, void* stream) {
    (void)stride_vec;
    if (nbatch <= 0) return 0;
    k_lanczos_init<<<nbatch, 1024, 0, (cudaStream_t)stream>>>(dphys, (const cplx*)x, stride_x, (cplx*)basis, stride_basis, nrm, active, kdone, el.ptr,
                                                               el.stride, er.ptr, er.stride, mask, n_zero);
    return done("yb_lanczos_init");
}

int yb_lanczos_step(int nbatch, int j, int kmax, double dt, double tol, yb_extent el, yb_extent er, int dphys, void* basis, int64_t stride_basis,
                    int64_t stride_vec, double* alpha, double* beta, const double* nrm, int32_t* active, int32_t* kdone, void* coef,
                    int32_t* n_active, int32_t* n_zero, void* stream) {
    if (nbatch <= 0) return 0;
    if (n_zero && n_zero == n_active) return fail_msg("yb_lanczos_step: n_zero must not alias n_active");
    if (kmax > 32 || j >= kmax) return fail_msg("yb_lanczos_step: kmax must be <= 32 and j < kmax");
    k_lanczos_step<<<nbatch, 1024, 0, (cudaStream_t)stream>>>(j, kmax, dt, tol, dphys, (cplx*)basis, stride_basis, stride_vec, alpha, beta, nrm, active,
                                                               kdone, (cplx*)coef, n_active, el.ptr, el.stride, er.ptr, er.stride, n_zero);
    return done("yb_lanczos_step");
}

int yb_lanczos_combine(int nbatch, const int32_t* mask, yb_extent el, yb_extent er, int dphys, const void* basis, int64_t stride_basis,
                       int64_t stride_vec, const void* coef, int kmax, const int32_t* kdone, void* out, int64_t stride_out, void* stream) {
    (void)kmax;
    if (nbatch <= 0) return 0;
    dim3 grid(grid_batch(stride_vec, 256, nbatch), nbatch);
    k_lanczos_combine<<<grid, 256, 0, (cudaStream_t)stream>>>(dphys, (const cplx*)basis, stride_basis, stride_vec, (const cplx*)coef, kdone, (cplx*)out,
                                                               stride_out, el.ptr, el.stride, er.ptr, er.stride, mask);
    return done("yb_lanczos_combine");
}

int yb_svd_prepare(int nbatch, const int32_t* mask, int mode, yb_extent el, yb_extent er, int dphys, const void* src, int64_t stride_src, void* w,
                   int64_t stride_w, void* q, int64_t stride_q, int32_t* meta, int32_t* rot_count, int32_t* n_zero, int32_t* status_zero,
                   void* stream) {
    if (nbatch <= 0) return 0;
    dim3 grid(grid_batch(stride_w, 256, nbatch), nbatch);
    k_svd_prepare<<<grid, 256, 0, (cudaStream_t)stream>>>(mode, dphys, (const cplx*)src, stride_src, (cplx*)w, stride_w, (cplx*)q, stride_q, meta,
                                                           rot_count, el.ptr, el.stride, er.ptr, er.stride, mask, n_zero, status_zero);
    return done("yb_svd_prepare");
}

int yb_svd_round(int nbatch, const int32_t* mask, int32_t* meta, int block_rows, int nblocks, int round, int full_panel, int inner_sweeps, void* w,
                 int64_t stride_w, void* q, int64_t stride_q, int32_t* rot_count, int max_len, int max_k, int32_t* n_zero, void* stream) {
    if (nbatch <= 0) return 0;
    if (nblocks < 2 || (nblocks & 1)) return fail_msg("yb_svd_round: nblocks must be even and >= 2");
    const size_t smem = (size_t)2 * block_rows * ((size_t)max_len + max_k) * sizeof(cplx);
    if (smem > 227 * 1024) return fail_msg("yb_svd_round: panel does not fit in shared memory");
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(k_svd_round, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail("yb_svd_round(attr)", e);
        configured = smem;
    }
    dim3 grid(nblocks / 2, nbatch);
    k_svd_round<<<grid, 512, smem, (cudaStream_t)stream>>>(block_rows, nblocks, round, full_panel, inner_sweeps, (cplx*)w, stride_w, (cplx*)q, stride_q,
                                                            meta, rot_count, max_len, max_k, mask, n_zero);
    return done("yb_svd_round");
}

int yb_svd_check(int nbatch, const int32_t* mask, int32_t* meta, int32_t* rot_count, int32_t* n_pending, void* stream) {
    if (nbatch <= 0) return 0;
    k_svd_check<<<(nbatch + 127) / 128, 128, 0, (cudaStream_t)stream>>>(nbatch, meta, rot_count, n_pending, mask);
    return done("yb_svd_check");
}

int yb_svd_finalize(int nbatch, const int32_t* mask, int32_t* meta, int trunc_mode, double threshold, int max_bond, int min_keep, int dist,
                    const void* w, int64_t stride_w, const void* q, int64_t stride_q, double* sval, int64_t stride_s, int32_t* keep,
                    int32_t keep_stride, void* uout, int64_t stride_u, void* vout, int64_t stride_v, int32_t* status, int max_k, void* stream) {
    if (nbatch <= 0) return 0;
    if (max_k > 1024) return fail_msg("yb_svd_finalize: max_k > 1024 unsupported");
    k_svd_finalize<<<nbatch, 512, 0, (cudaStream_t)stream>>>(trunc_mode, threshold, max_bond, min_keep, dist, (const cplx*)w, stride_w, (const cplx*)q,
                                                              stride_q, meta, sval, stride_s, keep, keep_stride, (cplx*)uout, stride_u, (cplx*)vout,
                                                              stride_v, status, mask);
    return done("yb_svd_finalize");
}

int yb_svd_finalize_top(int nbatch, const int32_t* mask, int32_t* meta, int trunc_mode, double threshold, int max_bond, int min_keep, int dist,
                        const void* w, int64_t stride_w, const void* q, int64_t stride_q, double* sval, int64_t stride_s, int32_t* keep,
                        int32_t keep_stride, void* uout, int64_t stride_u, void* vout, int64_t stride_v, int32_t* status, int max_k,
                        const double* lam, int64_t stride_lam, int top, void* stream) {
    if (nbatch <= 0) return 0;
    if (max_k > 1024) return fail_msg("yb_svd_finalize_top: max_k > 1024 unsupported");
    if (top <= 0 || !lam) return fail_msg("yb_svd_finalize_top: needs the sorted eigenvalues and top > 0");
    if (max_bond <= 0 || max_bond > top) return fail_msg("yb_svd_finalize_top: the split must be capped at max_bond <= top");
    k_svd_finalize<<<nbatch, 512, 0, (cudaStream_t)stream>>>(trunc_mode, threshold, max_bond, min_keep, dist, (const cplx*)w, stride_w, (const cplx*)q,
                                                              stride_q, meta, sval, stride_s, keep, keep_stride, (cplx*)uout, stride_u, (cplx*)vout,
                                                              stride_v, status, mask, lam, stride_lam, top);
    return done("yb_svd_finalize_top");
}

int yb_svd_ambiguous(int nbatch, const int32_t* mask, const int32_t* meta, const double* sval, int64_t stride_s, int trunc_mode, double threshold,
                     int two_level, double tau_rel, double noise_rel, const int32_t* status, int32_t* amb, int32_t* n_amb, void* stream) {
    if (nbatch <= 0) return 0;
    if (trunc_mode != 0 && trunc_mode != 1) return fail_msg("yb_svd_ambiguous: the Gram route serves truncation modes 0 and 1 only");
    k_svd_ambiguous<<<(nbatch + 127) / 128, 128, 0, (cudaStream_t)stream>>>(nbatch, meta, sval, stride_s, trunc_mode, threshold, two_level, tau_rel,
                                                                             noise_rel, status, amb, n_amb, mask);
    return done("yb_svd_ambiguous");
}

int yb_svd_select(int nbatch, const int32_t* mask, const int32_t* meta, const void* w, int64_t stride_w, double tau_rel, int32_t* sel_n,
                  int32_t* sel_idx, int max_k, void* stream) {
    if (nbatch <= 0) return 0;
    if (max_k > 1024) return fail_msg("yb_svd_select: max_k > 1024 unsupported");
    k_svd_select<<<nbatch, 512, 0, (cudaStream_t)stream>>>((const cplx*)w, stride_w, meta, tau_rel, sel_n, sel_idx, max_k, mask);
    return done("yb_svd_select");
}

int yb_svd_gather(int nbatch, const int32_t* mask, const int32_t* meta, const int32_t* sel_n, const int32_t* sel_idx, int max_k, int dir, void* w,
                  int64_t stride_w, void* q, int64_t stride_q, void* ws, int64_t stride_ws, void* qs, int64_t stride_qs, void* stream) {
    if (nbatch <= 0) return 0;
    dim3 grid(32, nbatch);
    k_svd_gather<<<grid, 256, 0, (cudaStream_t)stream>>>(dir, (cplx*)w, stride_w, (cplx*)q, stride_q, (cplx*)ws, stride_ws, (cplx*)qs, stride_qs, meta,
                                                          sel_n, sel_idx, max_k, mask);
    return done("yb_svd_gather");
}

int yb_heev(int nbatch, const int32_t* mask, yb_extent ek, void* g, int64_t stride_g, void* qh, int64_t stride_q, void* zt, int64_t stride_z,
            double* zr, int64_t stride_zr, double* d, double* e, void* tau, int64_t stride_vec, int32_t* meta, int32_t* status, int max_k,
            void* stream) {
    if (nbatch <= 0) return 0;
    if (max_k > heev::KMAX) return fail_msg("yb_heev: max_k > 512 unsupported");
    static bool configured = false;
    if (!configured) {
        cudaError_t e1 = cudaFuncSetAttribute(heev::k_hetrd<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)heev::SMEM_BYTES);
        if (e1 == cudaSuccess) e1 = cudaFuncSetAttribute(heev::k_ungtr<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)heev::SMEM_BYTES);
#define YB_RING_ATTR(R)                                                                                                                         \
        if (e1 == cudaSuccess) e1 = cudaFuncSetAttribute(heev::k_hetrd<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(heev::SMEM_BYTES + heev::ring_bytes(R))); \
        if (e1 == cudaSuccess) e1 = cudaFuncSetAttribute(heev::k_ungtr<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(heev::SMEM_BYTES + heev::ring_bytes(R)));
        YB_RING_ATTR(6) YB_RING_ATTR(12) YB_RING_ATTR(16)
#undef YB_RING_ATTR
        if (e1 != cudaSuccess) return fail("yb_heev(attr)", e1);
        e1 = cudaFuncSetAttribute(heev::k_stedc<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(stedc::Smem<128>));
        if (e1 == cudaSuccess) e1 = cudaFuncSetAttribute(heev::k_stedc<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(stedc::Smem<256>));
        if (e1 == cudaSuccess) e1 = cudaFuncSetAttribute(heev::k_stedc<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(stedc::Smem<512>));
        if (e1 != cudaSuccess) return fail("yb_heev(attr)", e1);
        configured = true;
    }
    cudaStream_t st = (cudaStream_t)stream;
    // k_ungtr (needs the reflectors) and k_steqr (needs d, e) are independent: the bandwidth-bound Q accumulation runs on a
    // side stream underneath the latency-bound QL chain.
    // (thread_local: every host thread that drives a lane of trajectories on its own stream gets its own side stream)
    static thread_local cudaStream_t side[16] = {nullptr};
    static thread_local cudaEvent_t ev_fork[16] = {nullptr}, ev_join[16] = {nullptr};
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 15;
    if (!side[dev]) {
        cudaError_t e2 = cudaStreamCreateWithFlags(&side[dev], cudaStreamNonBlocking);
        if (e2 == cudaSuccess) e2 = cudaEventCreateWithFlags(&ev_fork[dev], cudaEventDisableTiming);
        if (e2 == cudaSuccess) e2 = cudaEventCreateWithFlags(&ev_join[dev], cudaEventDisableTiming);
        if (e2 != cudaSuccess) return fail("yb_heev(side stream)", e2);
    }
    // row-ring depth of the streaming kernels: option "heev_ring" = -1 (default: by matrix size), 0 (direct loads), 6, 12, 16
    int ring = g_opt_heev_ring.load();
    // measured (profiles/r01_heev_microbench_v6_ring_depth.json): k = 512, B = 74: 59.3 -> 46.2 ms with 16 rows in flight; at
    // k = 256 the direct loads win for the lane batches in use (B = 128: 8.6 ms vs 9.1 ms)
    if (ring < 0) ring = (max_k > 256) ? 16 : 0;
    ring = ring >= 16 ? 16 : ring >= 12 ? 12 : ring >= 6 ? 6 : 0;
#define YB_HETRD(R) heev::k_hetrd<R><<<nbatch, heev::THREADS, heev::SMEM_BYTES + (R ? heev::ring_bytes(R) : 0), st>>>((cplx*)g, stride_g, ek.ptr, ek.stride, d, e, (cplx*)tau, stride_vec, mask)
    if (ring == 16) YB_HETRD(16); else if (ring == 12) YB_HETRD(12); else if (ring == 6) YB_HETRD(6); else YB_HETRD(0);
#undef YB_HETRD
    if (int rc = done("yb_heev(hetrd)")) return rc;
    cudaEventRecord(ev_fork[dev], st);
    cudaStreamWaitEvent(side[dev], ev_fork[dev], 0);
#define YB_UNGTR(R) heev::k_ungtr<R><<<nbatch, heev::THREADS, heev::SMEM_BYTES + (R ? heev::ring_bytes(R) : 0), side[dev]>>>((const cplx*)g, stride_g, (const cplx*)tau, stride_vec, (cplx*)qh, stride_q, ek.ptr, ek.stride, mask)
    if (ring == 16) YB_UNGTR(16); else if (ring == 12) YB_UNGTR(12); else if (ring == 6) YB_UNGTR(6); else YB_UNGTR(0);
#undef YB_UNGTR
    if (int rc = done("yb_heev(ungtr)")) return rc;
    cudaEventRecord(ev_join[dev], side[dev]);
    const bool use_ql = g_opt_heev_ql.load() != 0;      // on: the implicit-QL kernel instead of divide & conquer
    if (use_ql) {
        if (max_k <= 128)
            heev::k_steqr<128><<<nbatch, 128 + 32, 0, st>>>(d, e, stride_vec, zr, stride_zr, (cplx*)zt, stride_z, ek.ptr, ek.stride, meta, status, mask);
        else if (max_k <= 256)
            heev::k_steqr<256><<<nbatch, 256 + 32, 0, st>>>(d, e, stride_vec, zr, stride_zr, (cplx*)zt, stride_z, ek.ptr, ek.stride, meta, status, mask);
        else
            heev::k_steqr<512><<<nbatch, 512 + 32, 0, st>>>(d, e, stride_vec, zr, stride_zr, (cplx*)zt, stride_z, ek.ptr, ek.stride, meta, status, mask);
    } else if (max_k <= 128) {
        heev::k_stedc<128><<<nbatch, stedc::THREADS, sizeof(stedc::Smem<128>), st>>>(d, e, stride_vec, zr, stride_zr, (cplx*)zt, stride_z, ek.ptr,
                                                                                  ek.stride, meta, status, mask);
    } else if (max_k <= 256) {
        heev::k_stedc<256><<<nbatch, stedc::THREADS, sizeof(stedc::Smem<256>), st>>>(d, e, stride_vec, zr, stride_zr, (cplx*)zt, stride_z, ek.ptr,
                                                                                  ek.stride, meta, status, mask);
    } else {
        heev::k_stedc<512><<<nbatch, stedc::THREADS, sizeof(stedc::Smem<512>), st>>>(d, e, stride_vec, zr, stride_zr, (cplx*)zt, stride_z, ek.ptr,
                                                                                  ek.stride, meta, status, mask);
    }
    cudaStreamWaitEvent(st, ev_join[dev], 0);
    return done("yb_heev(stedc)");
}

// Round-2 eigen-solver: blocked tridiagonalisation, divide & conquer, eigenvector rows straight from the reflectors (yb_hetrd.cuh).
static int heev_rows_impl(int nbatch, const int32_t* mask, yb_extent ek, void* g, int64_t stride_g, void* eh, int64_t stride_eh, double* zr,
                          int64_t stride_zr, double* d, double* e, void* tau, int64_t stride_vec, int32_t* meta, int32_t* status, int max_k, int top,
                          int32_t* order, double* dsorted, int32_t* ntop, void* stream);

int yb_heev_rows(int nbatch, const int32_t* mask, yb_extent ek, void* g, int64_t stride_g, void* eh, int64_t stride_eh, double* zr, int64_t stride_zr,
                 double* d, double* e, void* tau, int64_t stride_vec, int32_t* meta, int32_t* status, int max_k, void* stream) {
    return heev_rows_impl(nbatch, mask, ek, g, stride_g, eh, stride_eh, zr, stride_zr, d, e, tau, stride_vec, meta, status, max_k, 0, nullptr, nullptr,
                          nullptr, stream);
}

int yb_heev_rows_top(int nbatch, const int32_t* mask, yb_extent ek, void* g, int64_t stride_g, void* eh, int64_t stride_eh, double* zr, int64_t stride_zr,
                     double* d, double* e, void* tau, int64_t stride_vec, int32_t* meta, int32_t* status, int max_k, int top, int32_t* order,
                     double* dsorted, int32_t* ntop, void* stream) {
    if (top <= 0 || !order || !dsorted || !ntop) return fail_msg("yb_heev_rows_top: needs top > 0 and the order / dsorted / ntop outputs");
    return heev_rows_impl(nbatch, mask, ek, g, stride_g, eh, stride_eh, zr, stride_zr, d, e, tau, stride_vec, meta, status, max_k, top, order, dsorted,
                          ntop, stream);
}

static int heev_rows_impl(int nbatch, const int32_t* mask, yb_extent ek, void* g, int64_t stride_g, void* eh, int64_t stride_eh, double* zr,
                          int64_t stride_zr, double* d, double* e, void* tau, int64_t stride_vec, int32_t* meta, int32_t* status, int max_k, int top,
                          int32_t* order, double* dsorted, int32_t* ntop, void* stream) {
    if (nbatch <= 0) return 0;
    const int expand = top > 0 ? 0 : 1;
    const double* srcR = top > 0 ? zr : nullptr;
    if (max_k > heev::KMAX) return fail_msg("yb_heev_rows: max_k > 512 unsupported");
    if (g == eh) return fail_msg("yb_heev_rows: the eigenvector rows must not alias the input matrix (it keeps the reflectors)");
    cudaStream_t st = (cudaStream_t)stream;
    static bool configured = false;
    if (!configured) {
        cudaError_t e1 = cudaFuncSetAttribute(hetrd::k_hetrd_blocked<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hetrd::smem_bytes<16>(256));
        if (e1 == cudaSuccess) e1 = cudaFuncSetAttribute(hetrd::k_hetrd_blocked<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hetrd::smem_bytes<8>(512));
        const int orm = (int)((size_t)(32 * 256 + 32) * sizeof(cplx));
        if (e1 == cudaSuccess) e1 = cudaFuncSetAttribute(hetrd::k_ormtr_rows<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, orm);
        if (e1 == cudaSuccess) e1 = cudaFuncSetAttribute(hetrd::k_ormtr_rows<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, orm);
        if (e1 == cudaSuccess) e1 = cudaFuncSetAttribute(hetrd::k_ormtr_rows<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, orm);
        if (e1 == cudaSuccess) e1 = cudaFuncSetAttribute(heev::k_stedc<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(stedc::Smem<128>));
        if (e1 == cudaSuccess) e1 = cudaFuncSetAttribute(heev::k_stedc<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(stedc::Smem<256>));
        if (e1 == cudaSuccess) e1 = cudaFuncSetAttribute(heev::k_stedc<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(stedc::Smem<512>));
        if (e1 != cudaSuccess) return fail("yb_heev_rows(attr)", e1);
        configured = true;
    }
    if (max_k <= 256)
        hetrd::k_hetrd_blocked<16><<<nbatch, hetrd::THREADS, hetrd::smem_bytes<16>(max_k), st>>>((cplx*)g, stride_g, ek.ptr, ek.stride, d, e, (cplx*)tau,
                                                                                              stride_vec, max_k, mask);
    else
        hetrd::k_hetrd_blocked<8><<<nbatch, hetrd::THREADS, hetrd::smem_bytes<8>(max_k), st>>>((cplx*)g, stride_g, ek.ptr, ek.stride, d, e, (cplx*)tau,
                                                                                            stride_vec, max_k, mask);
    if (int rc = done("yb_heev_rows(hetrd)")) return rc;
    if (max_k <= 128)
        heev::k_stedc<128><<<nbatch, stedc::THREADS, sizeof(stedc::Smem<128>), st>>>(d, e, stride_vec, zr, stride_zr, (cplx*)eh, stride_eh, ek.ptr, ek.stride,
                                                                                  meta, status, mask, expand);
    else if (max_k <= 256)
        heev::k_stedc<256><<<nbatch, stedc::THREADS, sizeof(stedc::Smem<256>), st>>>(d, e, stride_vec, zr, stride_zr, (cplx*)eh, stride_eh, ek.ptr, ek.stride,
                                                                                  meta, status, mask, expand);
    else
        heev::k_stedc<512><<<nbatch, stedc::THREADS, sizeof(stedc::Smem<512>), st>>>(d, e, stride_vec, zr, stride_zr, (cplx*)eh, stride_eh, ek.ptr, ek.stride,
                                                                                  meta, status, mask, expand);
    if (int rc = done("yb_heev_rows(stedc)")) return rc;
    if (top > 0) {
        heev::k_heev_order<<<nbatch, 512, 0, st>>>(d, stride_vec, ek.ptr, ek.stride, top, order, dsorted, ntop, mask);
        if (int rc = done("yb_heev_rows_top(order)")) return rc;
    }
    if (max_k <= 128) {
        const int pb = 32;
        hetrd::k_ormtr_rows<4><<<nbatch, hetrd::THREADS, (size_t)(pb * 128 + pb) * sizeof(cplx), st>>>((const cplx*)g, stride_g, (const cplx*)tau, stride_vec,
                                                                                                    (cplx*)eh, stride_eh, ek.ptr, ek.stride, pb, mask, srcR, stride_zr, order, stride_vec, top);
    } else if (max_k <= 256) {
        const int pb = 32;
        hetrd::k_ormtr_rows<8><<<nbatch, hetrd::THREADS, (size_t)(pb * 256 + pb) * sizeof(cplx), st>>>((const cplx*)g, stride_g, (const cplx*)tau, stride_vec,
                                                                                                    (cplx*)eh, stride_eh, ek.ptr, ek.stride, pb, mask, srcR, stride_zr, order, stride_vec, top);
    } else {
        const int pb = 16;
        hetrd::k_ormtr_rows<16><<<nbatch, hetrd::THREADS, (size_t)(pb * 512 + pb) * sizeof(cplx), st>>>((const cplx*)g, stride_g, (const cplx*)tau, stride_vec,
                                                                                                     (cplx*)eh, stride_eh, ek.ptr, ek.stride, pb, mask, srcR, stride_zr, order, stride_vec, top);
    }
    return done("yb_heev_rows(ormtr)");
}

int yb_cholqr_factor(int nbatch, const int32_t* mask, const int32_t* meta, const void* g, int64_t stride_g, void* linv, int64_t stride_l,
                     double threshold, int32_t* flag, int32_t* n_flag, int max_k, void* stream) {
    if (nbatch <= 0) return 0;
    if (max_k > cholqr::KMAX) return fail_msg("yb_cholqr_factor: max_k > 128 unsupported (packed triangle must fit in shared memory)");
    if (max_k < 1) max_k = 1;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(cholqr::k_cholqr_factor, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cholqr::smem_bytes(cholqr::KMAX));
        if (e != cudaSuccess) return fail("yb_cholqr_factor(attr)", e);
        configured = true;
    }
    cholqr::k_cholqr_factor<<<nbatch, cholqr::THREADS, cholqr::smem_bytes(max_k), (cudaStream_t)stream>>>((const cplx*)g, stride_g, (cplx*)linv, stride_l, meta,
                                                                                                         threshold, flag, n_flag, max_k, mask);
    return done("yb_cholqr_factor");
}

int yb_cholqr_refine(int nbatch, const int32_t* mask, const int32_t* meta, const void* g2, int64_t stride_g, void* x, int64_t stride_x, double tol,
                     int32_t* flag, int32_t* n_flag, void* stream) {
    if (nbatch <= 0) return 0;
    cholqr::k_cholqr_refine<<<nbatch, cholqr::THREADS, 0, (cudaStream_t)stream>>>((const cplx*)g2, stride_g, (cplx*)x, stride_x, meta, tol, flag, n_flag, mask);
    return done("yb_cholqr_refine");
}

int yb_cholqr_finalize(int nbatch, const int32_t* mask, int32_t* meta, const void* y2, int64_t stride_y, const void* l2, int64_t stride_l, int32_t* keep,
                       int32_t keep_stride, void* uout, int64_t stride_u, void* vout, int64_t stride_v, const int32_t* flag, int max_elems, void* stream) {
    if (nbatch <= 0) return 0;
    dim3 grid(grid_batch(max_elems, 256, nbatch), nbatch);
    cholqr::k_cholqr_finalize<<<grid, 256, 0, (cudaStream_t)stream>>>((const cplx*)y2, stride_y, (const cplx*)l2, stride_l, meta, keep, keep_stride, (cplx*)uout,
                                                                       stride_u, (cplx*)vout, stride_v, flag, mask);
    return done("yb_cholqr_finalize");
}

int yb_unpack_vh(int nbatch, const int32_t* mask, yb_extent ek, yb_extent er, int dphys, const void* src, int64_t stride_src, void* dst,
                 int64_t stride_dst, int max_elems, yb_extent set_ext, void* stream) {
    if (nbatch <= 0) return 0;
    dim3 grid(grid_batch(max_elems, 256, nbatch), nbatch);
    k_unpack_vh<<<grid, 256, 0, (cudaStream_t)stream>>>(dphys, (const cplx*)src, stride_src, (cplx*)dst, stride_dst, ek.ptr, ek.stride, er.ptr,
                                                         er.stride, mask, (int*)set_ext.ptr, set_ext.stride);
    return done("yb_unpack_vh");
}

int yb_copy_packed(int nbatch, const int32_t* mask, yb_extent e0, yb_extent e1, int mult, const void* src, int64_t stride_src, void* dst,
                   int64_t stride_dst, int max_elems, yb_extent set_ext, void* stream) {
    if (nbatch <= 0) return 0;
    dim3 grid(grid_batch(max_elems, 256, nbatch), nbatch);
    k_copy_packed<<<grid, 256, 0, (cudaStream_t)stream>>>(mult, (const cplx*)src, stride_src, (cplx*)dst, stride_dst, e0.ptr, e0.stride, e1.ptr,
                                                           e1.stride, mask, (int*)set_ext.ptr, set_ext.stride);
    return done("yb_copy_packed");
}

int yb_apply_local_op(int nbatch, const int32_t* mask, const int32_t* sel, int sel_value, yb_extent el, yb_extent er, int dphys, const void* op,
                      int64_t op_stride, void* a, int64_t stride_a, int max_lr, void* stream) {
    if (nbatch <= 0) return 0;
    if (dphys > 4) return fail_msg("yb_apply_local_op: dphys > 4 unsupported");
    dim3 grid(grid_batch(max_lr, 256, nbatch), nbatch);
    k_apply_local_op<<<grid, 256, 0, (cudaStream_t)stream>>>(dphys, (const cplx*)op, op_stride, (cplx*)a, stride_a, el.ptr, el.stride, er.ptr,
                                                              er.stride, mask, sel, sel_value);
    return done("yb_apply_local_op");
}

int yb_mpo_site_apply(int nbatch, const int32_t* mask, yb_extent el, yb_extent er, int dphys, int d_left, int d_right, const void* w,
                      const void* src, int64_t stride_src, void* dst, int64_t stride_dst, int max_out_lr, void* stream) {
    if (nbatch <= 0) return 0;
    if (dphys > 4) return fail_msg("yb_mpo_site_apply: dphys > 4 unsupported");
    if (d_left < 1 || d_right < 1) return fail_msg("yb_mpo_site_apply: MPO bond dimensions must be >= 1");
    if (src == dst) return fail_msg("yb_mpo_site_apply: source and destination must not alias");
    dim3 grid(grid_batch(max_out_lr, 256, nbatch), nbatch);
    k_mpo_site_apply<<<grid, 256, 0, (cudaStream_t)stream>>>(dphys, d_left, d_right, (const cplx*)w, (const cplx*)src, stride_src, (cplx*)dst,
                                                              stride_dst, el.ptr, el.stride, er.ptr, er.stride, mask);
    return done("yb_mpo_site_apply");
}

int yb_norm2(int nbatch, const int32_t* mask, yb_extent el, yb_extent er, int dphys, const void* a, int64_t stride_a, double* out, void* stream) {
    if (nbatch <= 0) return 0;
    k_norm2<<<nbatch, 1024, 0, (cudaStream_t)stream>>>(dphys, (const cplx*)a, stride_a, out, el.ptr, el.stride, er.ptr, er.stride, mask);
    return done("yb_norm2");
}

int yb_scale(int nbatch, const int32_t* mask, yb_extent el, yb_extent er, int dphys, const double* scale, void* a, int64_t stride_a, int max_elems,
             int rsqrt_mode, void* stream) {
    if (nbatch <= 0) return 0;
    dim3 grid(grid_batch(max_elems, 256, nbatch), nbatch);
    k_scale<<<grid, 256, 0, (cudaStream_t)stream>>>(dphys, scale, (cplx*)a, stride_a, el.ptr, el.stride, er.ptr, er.stride, mask, rsqrt_mode);
    return done("yb_scale");
}

int yb_local_rdm(int nbatch, const int32_t* mask, yb_extent el, yb_extent er, int dphys, const void* a, int64_t stride_a, const void* x,
                 int64_t stride_x, void* g, void* stream) {
    if (nbatch <= 0) return 0;
    if (dphys > 4) return fail_msg("yb_local_rdm: dphys > 4 unsupported");
    k_local_rdm<<<nbatch, 512, 0, (cudaStream_t)stream>>>(dphys, (const cplx*)a, stride_a, (const cplx*)x, stride_x, (cplx*)g, el.ptr, el.stride, er.ptr,
                                                           er.stride, mask);
    return done("yb_local_rdm");
}

int yb_philox_uniform(int nbatch, uint64_t seed, const int64_t* traj_index, uint32_t tag, int nblocks, uint32_t block0, double* out,
                      int64_t stride_out, void* stream) {
    if (nbatch <= 0 || nblocks <= 0) return 0;
    if (stride_out < 2LL * nblocks || (stride_out & 1) || ((uintptr_t)out & 15)) return fail_msg("yb_philox_uniform: rows must be 16-byte aligned and hold 2 * nblocks doubles");
    dim3 grid(grid_for(nblocks, 128, 64), nbatch);
    k_philox_uniform<<<grid, 128, 0, (cudaStream_t)stream>>>((unsigned long long)seed, (const long long*)traj_index, tag, nblocks, block0, out, stride_out);
    return done("yb_philox_uniform");
}

int yb_jump_test(int nbatch, const double* norm2, const double* table, int64_t table_stride, int64_t col, int64_t* pos, int use_philox,
                 uint64_t seed, const int64_t* traj_index, uint32_t tag, uint32_t block0, double* dp, double* u1, double* u2, int32_t* jmask,
                 int32_t* njump, void* stream) {
    if (nbatch <= 0) return 0;
    if (!use_philox && !table) return fail_msg("yb_jump_test: table mode needs the uniform table");
    if (use_philox && !traj_index) return fail_msg("yb_jump_test: philox mode needs the global trajectory indices");
    k_jump_test<<<1, 256, 0, (cudaStream_t)stream>>>(nbatch, norm2, table, table_stride, col, (long long*)pos, use_philox, (unsigned long long)seed,
                                                      (const long long*)traj_index, tag, block0, dp, u1, u2, jmask, njump);
    return done("yb_jump_test");
}

int yb_jump_select(int nbatch, const int32_t* jmask, const double* u2, const void* rdm1, const void* rdm2, int nproc, const int32_t* proc_site,
                   const int32_t* proc_two, const void* proc_ldl, const double* proc_strength, double dt, const void* ops_table, int nops, double* w,
                   double* total, int32_t* choice, void* ops_sel, int32_t* status, void* stream) {
    if (nbatch <= 0 || nproc <= 0) return 0;
    k_jump_select<<<(nbatch + 63) / 64, 64, 0, (cudaStream_t)stream>>>(nbatch, jmask, u2, (const cplx*)rdm1, (const cplx*)rdm2, nproc, proc_site, proc_two,
                                                                        (const cplx*)proc_ldl, proc_strength, dt, (const cplx*)ops_table, nops, w, total,
                                                                        choice, (cplx*)ops_sel, status);
    return done("yb_jump_select");
}

int yb_observables(int nbatch, int nobs, const int32_t* col, const int32_t* site, const int32_t* two, const void* mats, const void* rdm1,
                   const void* rdm2, double* out, int64_t out_stride, int32_t* status, void* stream) {
    if (nbatch <= 0 || nobs <= 0) return 0;
    const long long n = (long long)nbatch * nobs;
    k_observables<<<(int)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(nbatch, nobs, col, site, two, (const cplx*)mats, (const cplx*)rdm1,
                                                                             (const cplx*)rdm2, out, out_stride, status);
    return done("yb_observables");
}

int yb_diagnostics(int nbatch, int length, int dphys, const int32_t* bond, double* out, int64_t out_stride, void* stream) {
    if (nbatch <= 0) return 0;
    k_diagnostics<<<(nbatch + 127) / 128, 128, 0, (cudaStream_t)stream>>>(nbatch, length, dphys, bond, out, out_stride);
    return done("yb_diagnostics");
}

int yb_mask_andnot(int nbatch, const int32_t* mask, const int32_t* flag, int32_t* work, void* stream) {
    if (nbatch <= 0) return 0;
    k_mask_andnot<<<(nbatch + 127) / 128, 128, 0, (cudaStream_t)stream>>>(nbatch, mask, flag, work);
    return done("yb_mask_andnot");
}

// ---- CUDA graphs of fixed launch sequences (the gauge shifts of a chain walk: every extent is read on the device, so one captured
// sequence per (site, direction, policy) serves every step).  Capture is thread-local: other lanes keep launching and reading back.
int yb_graph_begin(void* stream) {
    cudaError_t e = cudaStreamBeginCapture((cudaStream_t)stream, cudaStreamCaptureModeThreadLocal);
    return e == cudaSuccess ? 0 : fail("yb_graph_begin", e);
}

struct YbGraph { cudaGraphExec_t exec; long long kernels; };      // the handle yb_graph_end returns: executable + its number of kernel nodes

int yb_graph_end(void* stream, void** graph_exec) {
    if (!graph_exec) return fail_msg("yb_graph_end: graph_exec is NULL");
    *graph_exec = nullptr;
    cudaGraph_t g = nullptr;
    cudaError_t e = cudaStreamEndCapture((cudaStream_t)stream, &g);
    if (e != cudaSuccess || !g) return fail("yb_graph_end(capture)", e != cudaSuccess ? e : cudaErrorUnknown);
    size_t nodes = 0;
    cudaGraphGetNodes(g, nullptr, &nodes);
    cudaGraphExec_t ex = nullptr;
    e = cudaGraphInstantiate(&ex, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) return fail("yb_graph_end(instantiate)", e);
    // the launches recorded during the capture were counted by done() although nothing ran: take them back, a replay counts them
    g_launches.fetch_sub((long long)nodes, std::memory_order_relaxed);
    *graph_exec = (void*)new YbGraph{ex, (long long)nodes};
    return 0;
}

int yb_graph_launch(void* graph_exec, void* stream) {
    if (!graph_exec) return fail_msg("yb_graph_launch: no graph");
    YbGraph* gr = (YbGraph*)graph_exec;
    cudaError_t e = cudaGraphLaunch(gr->exec, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail("yb_graph_launch", e);
    g_launches.fetch_add(gr->kernels, std::memory_order_relaxed);     // yb_launch_count counts KERNELS, replayed ones included
    return 0;
}

int yb_graph_destroy(void* graph_exec) {
    if (graph_exec) {
        YbGraph* gr = (YbGraph*)graph_exec;
        cudaGraphExecDestroy(gr->exec);
        delete gr;
    }
    return 0;
}

}  // extern "C"
