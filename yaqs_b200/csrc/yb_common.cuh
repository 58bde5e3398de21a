// yaqs_b200 -- shared device helpers and the batched-operation parameter block.
//
// Data model (see DESIGN.md "Data layout in HBM"):
//   * complex128 everywhere, interleaved (re, im) == double2, 16-byte aligned.
//   * A batch holds B trajectories.  Every per-trajectory tensor lives in a fixed-stride SLOT
//     (slot = capacity for the largest admissible bond dims) and is stored PACKED in C order with its
//     ACTUAL extents, exactly like the numpy arrays of the reference (MPS site tensor (s, l, r),
//     environment block (chi_ket, D, chi_bra), merged theta ((s s'), l, r)).
//   * The actual bond dimensions are device-resident: bond[b * (L+1) + j], j = 0..L, bond[.,0]=bond[.,L]=1.
//   * Optional per-trajectory int mask: kernels return immediately for b with mask[b] == 0.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

typedef double2 cplx;

__device__ __forceinline__ cplx cmul(cplx a, cplx b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cplx cfma(cplx a, cplx b, cplx c) {   // a*b + c
    return make_double2(fma(a.x, b.x, fma(-a.y, b.y, c.x)), fma(a.x, b.y, fma(a.y, b.x, c.y)));
}
__device__ __forceinline__ cplx cconj(cplx a) { return make_double2(a.x, -a.y); }
__device__ __forceinline__ double cabs2(cplx a) { return a.x * a.x + a.y * a.y; }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum of up to NV doubles per thread; result valid in ALL threads.  `red` needs 32*NV doubles.
template <int NV>
__device__ __forceinline__ void block_sum(double* v, double* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int q = 0; q < NV; ++q) v[q] = warp_sum(v[q]);
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int q = 0; q < NV; ++q) red[warp * NV + q] = v[q];
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < NV; ++q) {
        double s = 0.0;
        for (int w = 0; w < nw; ++w) s += red[w * NV + q];
        v[q] = s;
    }
}

// ---------------------------------------------------------------------------------------------
// GEMM op codes (the batched complex128 contraction family; every one is C = sum_seg A_seg * B_seg)
// ---------------------------------------------------------------------------------------------
enum YbGemmOp {
    YB_OP_RCONTRACT = 0,   // T1[(p,l),(a',r')] = sum_r X[(p,l),r] R[r,(a',r')]            primitives.py:201 / :105
    YB_OP_LCONTRACT = 1,   // Y_o[l',r']        = sum_(l,a) Lenv[(l,a),l'] T2_o[(l,a),r']    primitives.py:203
    YB_OP_ENVR3 = 2,       // R'[(l,a),l']      = sum_o sum_r' T2_o[(l,a),r'] conj(bra_o[l',r'])   primitives.py:108
    YB_OP_ENVL1 = 3,       // U1_o[(l,a),r']    = sum_l' Lenv[(l,a),l'] conj(bra_o[l',r'])         primitives.py:134
    YB_OP_ENVL3 = 4,       // L'[r,(a',r')]     = sum_(p,l) ket[(p,l),r] U2[(p,l),(a',r')]         primitives.py:136
    YB_OP_MERGE = 5,       // theta_(s s')[l,r] = sum_c A_s[l,c] B_s'[c,r]                          decompositions.py:100
    YB_OP_ABSORB_R = 6,    // next'_s[k,r]      = sum_j G[k,j] next_s[j,r]                          mps.py:742-746
    YB_OP_ABSORB_L = 7,    // prev'_s[l,k]      = sum_j prev_s[l,j] G[j,k]
    // Gram-matrix SVD path (yb_heev.cu): plain 2-D products on the packed Jacobi work matrix W (k rows of length len)
    YB_OP_GRAM = 8,        // G[i,j]            = sum_e W[i,e] conj(W[j,e])                         x0 = k, x1 = len
    YB_OP_EVEC = 9,        // Qacc[c,j]         = sum_i Zt[c,i] conj(Qh[j,i])   (rows = conj eigenvectors)   x0 = k
    YB_OP_ROTATE = 10,     // Wf[c,e]           = sum_j Qacc[c,j] W[j,e]                            x0 = k, x1 = len
    YB_OP_ROTATE_TOP = 11, // Wf[c,e]           = sum_j Etop[c,j] W[j,e], c < x2 (the eigenvector rows a capped split can keep)   x0 = k, x1 = len, x2 = rows
    YB_OP_LAST = 11,
};

struct YbGemmParams {
    int op;
    int nbatch;
    int ngroups;          // grid.y
    const int* mask;      // optional [B]
    // extents: every op reads up to three per-trajectory ints; e0/e1/e2 point at the FIRST trajectory's
    // value, consecutive trajectories are `estride` ints apart (bond arrays: estride = L+1; keep arrays: 1).
    const int* e0; const int* e1; const int* e2;
    int es0, es1, es2;
    int dphys;            // physical dimension of the tensor operand (2, 4, or 1 for bond tensors)
    int Dl, Dr;           // MPO bond dims left / right of the operand
    const cplx* A; const cplx* Bm; cplx* C;
    long long sA, sB, sC; // per-trajectory slot strides (elements)
    int conj_extra;       // != 0: conjugate the B operand in addition to the op's own rule
    int ragged;           // != 0: interleaved sub-tiles that skip padding (A/B option "gemm_ragged")
    int small_max;        // launches whose capacity M and N are both <= small_max take the element kernel (option "gemm_small_max", default 32)
    int no_small;         // != 0: never take the small-product kernel (option "gemm_no_small": A/B runs and the tile kernel's edge tests)
    int warps16;          // != 0: 16 warps with 16 x 32 warp tiles instead of 8 warps with 32 x 32 (option "gemm_warps16")
    unsigned long long* flops;   // optional device counter: every CTA adds the real flops of its tile (8 per complex MAC)
};
