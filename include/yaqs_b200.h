/* yaqs_b200 -- C ABI of the B200-native batched Tensor Jump Method kernels.
 *
 * Drop-in boundary: the reference (munich-quantum-toolkit/yaqs, pure Python) has no FFI; the
 * place a maintainer would bind these entry points is the per-trajectory backend boundary
 *   analog_tjm_1 / analog_tjm_2            src/mqt/yaqs/analog/analog_tjm.py:206,369
 *   run_backend_parallel(worker_fn=...)    src/mqt/yaqs/core/parallel_utils.py:331
 * with each entry point below replacing the numpy/scipy call(s) cited next to it (paths relative to
 * src/mqt/yaqs/core).  INTEGRATION.md shows the ctypes stub.
 *
 * Conventions
 *   - All tensors are complex128, interleaved (re, im), in device memory owned by the caller (torch CUDA tensors in
 *     the Python host); the library never allocates or frees tensor memory.
 *   - A batch holds `nbatch` trajectories.  Operand p of trajectory b starts at p + b * stride_p (strides in complex
 *     elements) and is stored packed, C order, with its ACTUAL extents (same index order as the reference's ndarrays).
 *   - Ragged bond dimensions are read on the device: an "extent" argument is a pointer to the first trajectory's int32
 *     plus the distance (in ints) to the next trajectory's value.
 *   - `mask` (optional, may be NULL): int32[nbatch]; trajectories with mask[b] == 0 are skipped.
 *   - Every function enqueues work on `stream` (a cudaStream_t passed as void*) and returns 0 on success, a negative
 *     value on error; yb_last_error() returns the message.  Nothing here synchronises the stream.
 */
#ifndef YAQS_B200_H
#define YAQS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { const int32_t* ptr; int32_t stride; } yb_extent;

const char* yb_last_error(void);
int yb_version(void);
/* Number of kernels this library has launched since the last call with reset != 0 (bench.py's gpu_launches). */
int64_t yb_launch_count(int reset);

/* Optional instrumentation: a device uint64 that every contraction CTA increments by the real flops of its tile (8 per
 * complex multiply-add); NULL (default) disables it.  bench.py uses it for roofline.achieved. */
void yb_set_flop_counter(void* device_u64);
/* Diagnostic switches for A/B runs (defaults = production route; environment YB_HEEV_QL / YB_HEEV_RING / YB_GEMM_RAGGED set the
 * initial value):
 *   "gemm_ragged" 1 = k_gemm with interleaved 8 x 8 sub-tiles that skip the padding of ragged tiles (bit-identical results;
 *               not yet measured on a B200)                                                                   (default 0)
 *   "wapply_onepass" 1 = k_wapply_onepass: every input element of the MPO-tensor step is read once and feeds all nx * ny outputs
 *               (bit-identical results; not yet measured on a B200; env YB_WAPPLY_ONEPASS)                     (default 0)
 *   "heev_ql"   1 = implicit-QL tridiagonal eigen-solver instead of divide & conquer                          (default 0)
 *   "heev_ring" depth of the cp.async row ring in k_hetrd / k_ungtr: 0 = direct register loads, 6, 12, 16;
 *               -1 = by matrix size (default: 16 rows in flight for k > 256, direct loads below)                      */
int yb_set_option(const char* name, int value);

/* Device memory a batch needs (SURVEY.md 8b "Ownership": the library never allocates; the caller sizes its buffers -- and the
 * number of trajectories per wave -- from this).  cap[0..length]: slot capacity of every bond (cap[0] = cap[length] = 1);
 * mpo_bond[0..length]: MPO bond dimensions; order: 1 or 2 (order 2 keeps the measurement copy psi next to phi); two_envs != 0: left and
 * right environments in separate storage (1-site / dynamic TDVP); gram != 0: include the Hermitian eigen-solver work matrices.
 * Pure host arithmetic: callable without a device. */
typedef struct {
    int64_t state, environments, theta, matvec, krylov, svd, heev, small, total;
} yb_workspace;
int yb_workspace_bytes(int length, int nbatch, int dphys, const int32_t* cap, const int32_t* mpo_bond, int order, int two_envs, int gram,
                       yb_workspace* out);

/* ---- contractions -------------------------------------------------------------------------------------------------
 * One grouped / K-segmented batched complex GEMM; `op` selects the index pattern (enum YbGemmOp in csrc/yb_common.cuh):
 *   0 RCONTRACT  T1[(p,l),(a',r')] = sum_r X[(p,l),r] R[r,(a',r')]          tensordot(ket, right_env, axes=1)     methods/tdvp/primitives.py:105,201
 *   1 LCONTRACT  Y_o[l',r'] = sum_(l,a) Lenv[(l,a),l'] T2_o[(l,a),r']       tensordot(t, left_env, ((2,1),(0,1))) primitives.py:203, project_bond :226
 *   2 ENVR3      R'[(l,a),l'] = sum_(o,r') T2_o[(l,a),r'] conj(bra_o[l',r'])  tensordot(t, bra.conj(), ...)        primitives.py:108
 *   3 ENVL1      U1_o[(l,a),r'] = sum_l' Lenv[(l,a),l'] conj(bra_o[l',r'])    tensordot(left_env, bra.conj(), (2,1)) primitives.py:134
 *   4 ENVL3      L'[r,(a',r')] = sum_(p,l) ket[(p,l),r] U2[(p,l),(a',r')]     tensordot(ket, t, ((0,1),(0,2)))     primitives.py:136
 *   5 MERGE      theta_(ss')[l,r] = sum_c A_s[l,c] B_s'[c,r]                  merge_two_site                       methods/decompositions.py:87-102
 *   6 ABSORB_R   next'_s[k,r] = sum_j G[k,j] next_s[j,r]                      R into the right neighbour          data_structures/mps.py:742-746
 *   7 ABSORB_L   prev'_s[l,k] = sum_j prev_s[l,j] G[j,k]                      (mirror image, left shifts)          mps.py:771-788
 *   8 GRAM       G[i,j] = sum_e W[i,e] conj(W[j,e])       Gram matrix of the packed SVD work matrix (k rows of length len)
 *   9 EVEC       Qacc[c,j] = sum_i Zt[c,i] conj(Qh[j,i])  conjugated eigenvectors of G as rows (yb_heev outputs)
 *  10 ROTATE     Wf[c,e] = sum_j Qacc[c,j] W[j,e]         rows of Wf are orthogonal with norms = singular values
 *  11 ROTATE_TOP Wf[c,e] = sum_j Etop[c,j] W[j,e], c < e2  the same product for the compact eigenvector block of yb_heev_rows_top (e0 = k, e1 = len, e2 = rows)
 * e0,e1,e2: the op's extents (l, r | l, c, r | k, j, r | l, j, k | k, len).  conj_b != 0 conjugates the B operand in addition.
 * max_m / max_n bound the per-trajectory GEMM M and N (grid sizing). */
int yb_gemm(int op, int nbatch, int ngroups, const int32_t* mask, yb_extent e0, yb_extent e1, yb_extent e2, int dphys, int d_left,
            int d_right, const void* a, int64_t stride_a, const void* b, int64_t stride_b, void* c, int64_t stride_c, int conj_b,
            int max_m, int max_n, void* stream);

/* project_site, first two contractions in one kernel (primitives.py:201-202; update_right_environment :105-106):
 *   out[x,l,y,r'] = sum_(u,v) wt[x,u,y,v] sum_r x[u,l,r] renv[r,v,r']
 * x: packed (dphys, l, r) local tensor; renv: (r, d_right, r) right environment block; wt: (dphys, dphys, d_left, d_right) MPO tensor
 * (W1 or the merged W2), shared by all trajectories; out: packed (dphys, l, d_left, r) -- what yb_gemm(LCONTRACT / ENVR3) consumes.
 * The MPO step runs in the GEMM epilogue while the tile is on chip (no intermediate in HBM).  dphys in {1, 2, 4}, d_right <= 16. */
int yb_rcontract_mpo(int nbatch, const int32_t* mask, yb_extent el, yb_extent er, int dphys, int d_left, int d_right, const void* x,
                     int64_t stride_x, const void* renv, int64_t stride_r, const void* wt, void* out, int64_t stride_out, int max_l, int max_r,
                     void* stream);

/* out[x,l,y,r] = sum_(u,v) wt[x,u,y,v] in[u,l,v,r]: the MPO-tensor step of project_site / update_*_environment
 * (tensordot(op, tensor, axes=((1,3),(0,2))) primitives.py:106,202; tensordot(op, tensor, ((0,2),(2,1))) :135 with the
 * host-transposed MPO tensor).  wt is shared by all trajectories. */
int yb_wapply(int nbatch, const int32_t* mask, yb_extent el, yb_extent er, int nx, int nu, int ny, int nv, const void* wt,
              const void* in, int64_t stride_in, void* out, int64_t stride_out, int max_lr, void* stream);

/* env[b][i,a,j] = delta_ij for a < D, i,j < chi[b]  (primitives.py:162-168, integrators.py:186-193). */
int yb_env_identity(int nbatch, const int32_t* mask, yb_extent echi, int dmpo, void* env, int64_t stride_env, int max_chi, void* stream);

/* ---- Lanczos / Krylov exponential (methods/matrix_exponential.py:33-201) ------------------------------------------
 * basis: [nbatch][kmax+1][nloc_max] Lanczos vectors; the matvec writes H v_j into slot j+1, yb_lanczos_step turns it into
 * v_{j+1} in place.  n_loc[b] = dphys * el[b] * er[b]. */
int yb_lanczos_init(int nbatch, const int32_t* mask, yb_extent el, yb_extent er, int dphys, const void* x, int64_t stride_x,
                    void* basis, int64_t stride_basis, int64_t stride_vec, double* nrm, int32_t* active, int32_t* kdone, int32_t* n_zero,
                    void* stream);
/* One Lanczos iteration for all active trajectories: alpha_j, orthogonalise, beta_j, breakdown / convergence test
 * (k x k tridiagonal eigen-solve = scipy.linalg.eigh_tridiagonal, linalg/eigh.py:29), normalise.  On finish writes the
 * k complex combination coefficients y (already times |v|) to coef[b][0..k) and kdone[b] = k, clears active[b].
 * n_active (device int32, zero on entry) receives the number of trajectories still iterating.  n_zero (optional, must not alias
 * n_active) is set to 0 by the same launch: with a PAIR of counters -- step j counts into c[j & 1] and clears c[(j + 1) & 1], and
 * yb_lanczos_init clears c[0] through its own n_zero -- the iteration needs no fill launch between its steps. */
int yb_lanczos_step(int nbatch, int j, int kmax, double dt, double tol, yb_extent el, yb_extent er, int dphys, void* basis,
                    int64_t stride_basis, int64_t stride_vec, double* alpha, double* beta, const double* nrm, int32_t* active,
                    int32_t* kdone, void* coef, int32_t* n_active, int32_t* n_zero, void* stream);
/* out = sum_j coef_j v_j for trajectories with kdone[b] > 0 (zero-norm inputs are returned untouched, :67-69). */
int yb_lanczos_combine(int nbatch, const int32_t* mask, yb_extent el, yb_extent er, int dphys, const void* basis, int64_t stride_basis,
                       int64_t stride_vec, const void* coef, int kmax, const int32_t* kdone, void* out, int64_t stride_out, void* stream);

/* ---- truncated SVD by one-sided Jacobi (linalg/svd.py:51 + linalg/svd_utils.py:22 + decompositions.py:105-185) ----
 * The matrix of trajectory b is M[b] (rows x cols), gathered from `src` by `mode`:
 *   0 THETA   src = merged theta ((s s'), l, r): rows (s, l), cols (s', r)           split_two_site :147-154
 *   1 ROWS    src = site tensor (s, l, r) read as (s l) x r                          right shift (mps.py:736-764)
 *   2 COLS    src = site tensor (s, l, r) read as l x (s r)                          left shift  (mps.py:771-788)
 * yb_svd_prepare builds the Jacobi work matrix W (the smaller side's vectors as contiguous rows) and Q = I.
 * yb_svd_round does one round of a cyclic block sweep (round r of `nrounds`), counting rotations per trajectory.
 * yb_svd_finalize sorts the singular values, applies truncate(mode, threshold, max_bond, min_keep) and writes
 *   keep[b], sval[b][:], U f(S) as rows x keep (row-major) to `uout`, g(S) Vh as keep x cols to `vout`
 *   (dist: 0 = S to the left factor, 1 = right, 2 = sqrt on both).
 * meta[b][7] carries the matrix scale (biased exponent of max |W_ij|^2, accumulated by yb_svd_prepare, cleared by
 *   yb_svd_finalize; the buffer must be zero before the first yb_svd_prepare): rows whose norm falls below tol * 2^(scale/2) are
 *   rounding residue of a rank-deficient input, take no part in further rotations and count as exact zeros in finalize. */
int yb_svd_prepare(int nbatch, const int32_t* mask, int mode, yb_extent el, yb_extent er, int dphys, const void* src, int64_t stride_src,
                   void* w, int64_t stride_w, void* q, int64_t stride_q, int32_t* meta, int32_t* rot_count, int32_t* n_zero, int32_t* status_zero,
                   void* stream);
int yb_svd_round(int nbatch, const int32_t* mask, int32_t* meta, int block_rows, int nblocks, int round, int full_panel, int inner_sweeps,
                 void* w, int64_t stride_w, void* q, int64_t stride_q, int32_t* rot_count, int max_len, int max_k, int32_t* n_zero, void* stream);
/* status_zero (optional, yb_svd_prepare): int32[nbatch] cleared by the launch (the scratch status words of the Gram route).
 * n_zero (optional, both calls): a device int32 cleared by the launch -- the counter yb_svd_check / yb_svd_ambiguous / yb_cholqr_* add
 * to afterwards, so a factorisation needs no separate fill launch. */
/* yb_svd_prepare `mode` may carry an orientation request in bits 4-5 (0: by shape, 1: W = M, 2: W = M^H).
 *
 * Gram-matrix route for the large matrices (replaces the bulk of zgesdd at linalg/svd.py:89-104): after yb_svd_prepare,
 *   yb_gemm(GRAM) -> yb_heev -> yb_gemm(EVEC) -> yb_gemm(ROTATE) -> yb_svd_finalize(w = Wf, q = Qacc).
 * yb_heev: Hermitian eigen-decomposition G = (Qh Z) diag(d) (Qh Z)^H of the k x k matrices in `g` (k = ek[b]; g is
 * overwritten with the Householder vectors): Householder tridiagonalisation, explicit Qh, implicit-QL with accumulated
 * rotations.  Outputs: qh (k x k), zt (k x k complex, row c = eigenvector c of the tridiagonal), d (eigenvalues, unsorted);
 * zr (k*k doubles), e, tau (k) are workspace; d/e/tau rows are stride_vec apart.  meta (optional, the SVD meta block) gets
 * done = 1 on convergence; status bit 1 is set otherwise. */
int yb_heev(int nbatch, const int32_t* mask, yb_extent ek, void* g, int64_t stride_g, void* qh, int64_t stride_q, void* zt, int64_t stride_z,
            double* zr, int64_t stride_zr, double* d, double* e, void* tau, int64_t stride_vec, int32_t* meta, int32_t* status, int max_k,
            void* stream);
/* yb_heev_rows (round 2): the same eigen-decomposition with the eigenvectors delivered as ROWS, eh[c][:] = (eigenvector c)^H -- what
 * yb_gemm(ROTATE) takes -- so the Gram route is yb_gemm(GRAM) -> yb_heev_rows -> yb_gemm(ROTATE) -> yb_svd_finalize.  Blocked
 * tridiagonalisation (zlatrd panels held in shared memory, one rank-2nb update of the trailing block per panel), divide & conquer on
 * the tridiagonal, then the reflectors are applied to the rows of Z^T directly (no explicit Qh, no EVEC product).  g is overwritten
 * with the reflectors (row j = v_j), eh must not alias g; zr (k*k doubles), e, tau are workspace; d = eigenvalues (unsorted). */
int yb_heev_rows(int nbatch, const int32_t* mask, yb_extent ek, void* g, int64_t stride_g, void* eh, int64_t stride_eh, double* zr, int64_t stride_zr,
                 double* d, double* e, void* tau, int64_t stride_vec, int32_t* meta, int32_t* status, int max_k, void* stream);
/* yb_heev_rows_top: yb_heev_rows for a split that is capped at max_bond <= top < k (split_two_site under max_bond_dim,
 * core/methods/decompositions.py:105-185 with svd_utils.py:99-101): only the eigenvectors of the `top` largest eigenvalues can be kept,
 * so only those rows are back-transformed.  order[b][rank] = index of the rank-th largest eigenvalue, dsorted[b][rank] = its value,
 * ntop[b] = min(k, top); eh rows 0 .. ntop[b]) = the eigenvector rows in that order (compact).  Follow with yb_gemm(ROTATE_TOP, e2 =
 * ntop) and yb_svd_finalize_top, which takes the singular values behind row ntop -- needed for the discarded weight only -- from
 * sqrt(max(dsorted, 0)). */
int yb_heev_rows_top(int nbatch, const int32_t* mask, yb_extent ek, void* g, int64_t stride_g, void* eh, int64_t stride_eh, double* zr, int64_t stride_zr,
                     double* d, double* e, void* tau, int64_t stride_vec, int32_t* meta, int32_t* status, int max_k, int top, int32_t* order,
                     double* dsorted, int32_t* ntop, void* stream);
int yb_svd_finalize_top(int nbatch, const int32_t* mask, int32_t* meta, int trunc_mode, double threshold, int max_bond, int min_keep, int dist,
                        const void* w, int64_t stride_w, const void* q, int64_t stride_q, double* sval, int64_t stride_s, int32_t* keep,
                        int32_t keep_stride, void* uout, int64_t stride_u, void* vout, int64_t stride_v, int32_t* status, int max_k,
                        const double* lam, int64_t stride_lam, int top, void* stream);
/* Second level of the Gram route: yb_svd_select lists, per trajectory, the rows of the rotated work matrix whose squared
 * norm is below tau_rel * max (sel_n[b] rows, indices in sel_idx[b][0..max_k)); yb_svd_gather copies those rows of (w, q)
 * into the compact blocks (ws, qs) (dir = 0) or back (dir = 1).  Between the two the caller re-diagonalises the block with
 * GRAM / yb_heev / EVEC / ROTATE using sel_n as the row extent. */
int yb_svd_select(int nbatch, const int32_t* mask, const int32_t* meta, const void* w, int64_t stride_w, double tau_rel, int32_t* sel_n,
                  int32_t* sel_idx, int max_k, void* stream);
int yb_svd_gather(int nbatch, const int32_t* mask, const int32_t* meta, const int32_t* sel_n, const int32_t* sel_idx, int max_k, int dir, void* w,
                  int64_t stride_w, void* q, int64_t stride_q, void* ws, int64_t stride_ws, void* qs, int64_t stride_qs, void* stream);
/* After a sweep: meta.done = (rot_count == 0); n_pending (device int32, zeroed by caller) counts unfinished trajectories. */
int yb_svd_check(int nbatch, const int32_t* mask, int32_t* meta, int32_t* rot_count, int32_t* n_pending, void* stream);
/* Gram route, closing test: amb[b] = 1 when trajectory b's keep decision lies within the error bound of the threshold (noise_rel *
 * s_max^2; tau_rel times less for second-level rows), its eigen-solve did not converge (meta done flag / status[b]) or its spectrum is
 * not finite -- those trajectories are redone by the one-sided Jacobi.  n_amb (one int, zeroed by the caller) counts them.  Replaces the
 * host-side comparison the reference never needs because LAPACK's zgesdd is backward stable (core/linalg/svd.py:89-104). */
int yb_svd_ambiguous(int nbatch, const int32_t* mask, const int32_t* meta, const double* sval, int64_t stride_s, int trunc_mode, double threshold,
                     int two_level, double tau_rel, double noise_rel, const int32_t* status, int32_t* amb, int32_t* n_amb, void* stream);
/* ---- truncation-free gauge shifts by Cholesky-QR (csrc/yb_cholqr.cuh) -----------------------------------------------------------
 * Replaces scipy.linalg.svd + truncate at core/data_structures/mps.py:747-764 (SVD gauge shifts with the hard-coded 1e-12 discarded
 * weight; the QR moves of :736-746) for the trajectories that provably keep EVERY singular value: truncate() (core/linalg/svd_utils.py:
 * 60-75) keeps all of them iff sigma_min^2 >= threshold, and then any orthonormal basis times the complementary factor is the same MPS
 * up to a unitary on the bond.  Sequence: yb_svd_prepare (orientation forced so that the bond side indexes the rows of w) ->
 * yb_gemm(GRAM: g = w w^H) -> yb_cholqr_factor -> yb_gemm(ROTATE: y1 = linv w) -> yb_gemm(GRAM: g2 = y1 y1^H) -> yb_cholqr_refine ->
 * yb_gemm(ROTATE: y2 = x y1) -> yb_gemm(GRAM: l2 = w y2^H) -> yb_cholqr_finalize.
 * yb_cholqr_factor: g = L L^H (packed lower triangle in shared memory, max_k <= 128), linv = L^-1 (k x k row-major, zeros above the
 * diagonal) and the certificate lambda_min(g) >= 1 / ||L^-1||_F^2: flag[b] = 0 iff that bound, less 64 k eps trace(g), is positive and
 * >= 4 * threshold; flag[b] = 1 otherwise (non-positive pivot, k > len, non-finite input, bound too small) -- the caller
 * redoes flagged trajectories with the eigen-solver route.  n_flag (one int, zeroed by the caller) counts them.
 * yb_cholqr_refine: x = (3 I - g2) / 2 (one Newton-Schulz step towards g2^-1/2); flags trajectories with ||g2 - I||_F > tol.
 * yb_cholqr_finalize: for unflagged trajectories keep = k and (uout, vout) in the layout of yb_svd_finalize: orient 1 (w = M^H):
 * U = y2^H, "S Vh" = l2^H; orient 0 (w = M): "U S" = l2, Vh = y2. */
int yb_cholqr_factor(int nbatch, const int32_t* mask, const int32_t* meta, const void* g, int64_t stride_g, void* linv, int64_t stride_l,
                     double threshold, int32_t* flag, int32_t* n_flag, int max_k, void* stream);
int yb_cholqr_refine(int nbatch, const int32_t* mask, const int32_t* meta, const void* g2, int64_t stride_g, void* x, int64_t stride_x, double tol,
                     int32_t* flag, int32_t* n_flag, void* stream);
int yb_cholqr_finalize(int nbatch, const int32_t* mask, int32_t* meta, const void* y2, int64_t stride_y, const void* l2, int64_t stride_l, int32_t* keep,
                       int32_t keep_stride, void* uout, int64_t stride_u, void* vout, int64_t stride_v, const int32_t* flag, int max_elems, void* stream);
int yb_svd_finalize(int nbatch, const int32_t* mask, int32_t* meta, int trunc_mode, double threshold, int max_bond, int min_keep,
                    int dist, const void* w, int64_t stride_w, const void* q, int64_t stride_q, double* sval, int64_t stride_s,
                    int32_t* keep, int32_t keep_stride, void* uout, int64_t stride_u, void* vout, int64_t stride_v, int32_t* status,
                    int max_k, void* stream);

/* ---- site-tensor utilities -----------------------------------------------------------------------------------------
 * dst(s', k, r) = src(k, (s', r)): unpack g(S) Vh rows into an MPS site tensor (decompositions.py:171-184). */
int yb_unpack_vh(int nbatch, const int32_t* mask, yb_extent ek, yb_extent er, int dphys, const void* src, int64_t stride_src, void* dst,
                 int64_t stride_dst, int max_elems, yb_extent set_ext, void* stream);
/* Packed copy of n[b] = mult * e0[b] * e1[b] complex elements per trajectory.
 * set_ext (both calls; ptr may be NULL): after the copy, trajectory b's int at set_ext is set to ek[b] / e0[b] -- the bond dimension
 * the factorisation kept is written into the state's bond array by the launch that installs the new site tensor (the array entry
 * must not be an extent of the same launch). */
int yb_copy_packed(int nbatch, const int32_t* mask, yb_extent e0, yb_extent e1, int mult, const void* src, int64_t stride_src, void* dst,
                   int64_t stride_dst, int max_elems, yb_extent set_ext, void* stream);
/* A_s <- sum_t op[s,t] A_t  with a dphys x dphys complex matrix, either shared (op_stride = 0) or per trajectory; optional
 * per-trajectory site selection: trajectory b is touched only if sel == NULL or sel[b] == sel_value.
 * (dissipation.py:125-126, stochastic_process.py:255,262-263, digital_tjm.py:304-309). */
int yb_apply_local_op(int nbatch, const int32_t* mask, const int32_t* sel, int sel_value, yb_extent el, yb_extent er, int dphys,
                      const void* op, int64_t op_stride, void* a, int64_t stride_a, int max_lr, void* stream);
/* dst[s, (l a), (r b)] = sum_t w[s, t, a, b] src[t, l, r]: one site of MPO.multiply on an MPS (mpo.py:1511-1537,
 * mpo_utils.py:27-56, the gate-MPO route of digital_tjm.py:536-556).  src is the packed (dphys, l, r) site tensor with l = el[b],
 * r = er[b]; dst receives the packed (dphys, l * d_left, r * d_right) tensor (MPS bond index major) and must not alias src; w is the
 * (dphys, dphys, d_left, d_right) MPO site tensor shared by all trajectories.  max_out_lr bounds l * d_left * r * d_right. */
int yb_mpo_site_apply(int nbatch, const int32_t* mask, yb_extent el, yb_extent er, int dphys, int d_left, int d_right, const void* w,
                      const void* src, int64_t stride_src, void* dst, int64_t stride_dst, int max_out_lr, void* stream);
/* out[b] = sum |a|^2 over the packed tensor (MPS.norm at the centre site, mps.py:1539-1565). */
int yb_norm2(int nbatch, const int32_t* mask, yb_extent el, yb_extent er, int dphys, const void* a, int64_t stride_a, double* out,
             void* stream);
/* a <- a * scale[b] (real); rsqrt_mode != 0: a <- a / sqrt(scale[b]) (scale = the squared norm of yb_norm2: mps.py:736-746 in one launch). */
int yb_scale(int nbatch, const int32_t* mask, yb_extent el, yb_extent er, int dphys, const double* scale, void* a, int64_t stride_a,
             int max_elems, int rsqrt_mode, void* stream);
/* g[b][o][p] = sum_(l,r) conj(a_o[l,r]) x_p[l,r]: local reduced density matrix used for observables
 * (mps.py:961-1047) and jump weights (stochastic_process.py:139-176). */
int yb_local_rdm(int nbatch, const int32_t* mask, yb_extent el, yb_extent er, int dphys, const void* a, int64_t stride_a, const void* x,
                 int64_t stride_x, void* g, void* stream);


/* ---- device RNG (rng_mode = "philox") ------------------------------------------------------------------------------
 * Counter-based replacement of the reference's per-trajectory numpy streams (core/random_utils.py:20-87): stream
 * (seed, traj, tag) with tag 0x5452414A ("TRAJ": initialize / step_through) or 0x53414D50 ("SAMP": block index = time step
 * of the measurement copy).  Philox4x32-10, key = seed, counter = (block0 + c, tag, traj_lo, traj_hi); block c yields
 * out[b][2c], out[b][2c+1] = the two uniforms one jump test can consume (methods/stochastic_process.py:190-292).
 * traj_index: device int64[nbatch] GLOBAL trajectory indices (tables do not depend on batching or on the GPU count).
 * Rows of `out` are stride_out doubles apart (even, >= 2 * nblocks, 16-byte aligned). */
int yb_philox_uniform(int nbatch, uint64_t seed, const int64_t* traj_index, uint32_t tag, int nblocks, uint32_t block0, double* out,
                      int64_t stride_out, void* stream);

/* ---- stochastic_process on the device (methods/stochastic_process.py:190-292) -------------------------------------------
 * yb_jump_test: dp[b] = 1 - norm2[b] (calculate_stochastic_factor, :36-60), jmask[b] = (u1 < dp) (:238), njump = number of jumps.
 * The two uniforms of a test sit at positions p, p + 1 of the trajectory's stream, p = pos[b] (or `col` when pos == NULL):
 * table mode (use_philox = 0): table[b * table_stride + p] -- streams replayed from the reference's seeded numpy generators;
 * philox mode: drawn in the kernel, component (q & 1) of Philox block (q >> 1), q = 2 * block0 + p, of stream (seed, traj_index[b],
 * tag) -- the same numbers yb_philox_uniform would put into a table.  pos[b] advances by 1 + jumped (what the reference's Generator
 * hands out); u1 / u2 return the uniforms (u2 feeds yb_jump_select and the test logs).
 * yb_jump_select (trajectories with jmask != 0): w[b][k] = dt * proc_strength[k] * Re sum_e proc_ldl[k][e] * rdm[site_k][b][e] with the
 * local reduced density matrices of the transfer sweep (rdm1: [L][B][4]; rdm2: [L - 1][B][16] for proc_two[k] != 0) -- :139-176;
 * total = sum_k w (zero / non-finite: status bit 4, the reference's ValueError :179-187); w is overwritten with the pdf w / total;
 * choice = #{k : cumsum(pdf)_k / cumsum(pdf)_last <= u2}, clamped (numpy's Generator.choice, :257); ops_sel[b] = ops_table[choice]
 * (nops complex numbers per process: the operator rows the application launches take). */
int yb_jump_test(int nbatch, const double* norm2, const double* table, int64_t table_stride, int64_t col, int64_t* pos, int use_philox,
                 uint64_t seed, const int64_t* traj_index, uint32_t tag, uint32_t block0, double* dp, double* u1, double* u2, int32_t* jmask,
                 int32_t* njump, void* stream);
int yb_jump_select(int nbatch, const int32_t* jmask, const double* u2, const void* rdm1, const void* rdm2, int nproc, const int32_t* proc_site,
                   const int32_t* proc_two, const void* proc_ldl, const double* proc_strength, double dt, const void* ops_table, int nops, double* w,
                   double* total, int32_t* choice, void* ops_sel, int32_t* status, void* stream);

/* ---- observables and diagnostics of a recorded column ------------------------------------------------------------------
 * yb_observables: out[b * out_stride + col[k]] = Re sum_e mats[k][e] * rdm[site[k]][b][e] for the nobs local observables (one-site:
 * 4 entries of rdm1 [L][B][4]; two[k] != 0: 16 entries of rdm2 [L - 1][B][16]) -- MPS.expect / evaluate_observables
 * (core/data_structures/mps.py:1178-1234) on the local reduced density matrices of the transfer sweep; an imaginary part >= 1e-13
 * raises status bit 8 (the reference asserts, :1233).
 * yb_diagnostics: out[b] = (sum chi^3, max(dphys, max chi), sum chi) over the inner bonds (mps.py:549-602), bond = int32 [B][L + 1]. */
int yb_observables(int nbatch, int nobs, const int32_t* col, const int32_t* site, const int32_t* two, const void* mats, const void* rdm1,
                   const void* rdm2, double* out, int64_t out_stride, int32_t* status, void* stream);
int yb_diagnostics(int nbatch, int length, int dphys, const int32_t* bond, double* out, int64_t out_stride, void* stream);

/* work[b] = (mask == NULL || mask[b]) && !flag[b]. */
int yb_mask_andnot(int nbatch, const int32_t* mask, const int32_t* flag, int32_t* work, void* stream);

/* ---- CUDA graphs of fixed launch sequences ------------------------------------------------------------------------------
 * The launches of one gauge shift (yb_svd_prepare ... yb_cholqr_finalize, the copies that install the factors) depend on the host only
 * through pointers and capacities that are fixed per (site, direction, policy); every extent is read on the device.  yb_graph_begin
 * puts `stream` into (thread-local) capture, the yb_* calls issued on it are recorded instead of executed, yb_graph_end returns an
 * executable graph, yb_graph_launch replays it on any stream: one host call per shift instead of ~14 (dispatch row of SURVEY 8b; the
 * reference has no counterpart -- it is one Python interpreter per trajectory, core/parallel_utils.py:331-390).  Kernels with
 * one-time attribute configuration must have run once before they are captured. */
int yb_graph_begin(void* stream);
int yb_graph_end(void* stream, void** graph_exec);
int yb_graph_launch(void* graph_exec, void* stream);
int yb_graph_destroy(void* graph_exec);

#ifdef __cplusplus
}
#endif
#endif /* YAQS_B200_H */
