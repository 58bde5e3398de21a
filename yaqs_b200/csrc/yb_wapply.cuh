// One-pass variant of the MPO-tensor step  out[x,l,y,r] = sum_(u,v) wt[x,u,y,v] in[u,l,v,r]  (primitives.py:106,135,202).
//
// k_wapply (yb_ops.cu) assigns one (x, y) output slab to a block, so every input slab (u, v) is read once per non-zero wt[x,u,y,v]:
// about 0.3 * nx * ny times the input (3.6x for a two-site Ising step).  Here a thread owns one (l, r) position, reads each input
// element ONCE and feeds it into all nx * ny accumulators (registers; compile-time bounds NX x NY with run-time extents), so the
// traffic is the algorithmic minimum: nu*nv reads + nx*ny writes per position, all coalesced along r.  The (u, v) loop order and the
// zero skip are those of k_wapply, so every output element is the same chain of fused multiply-adds: results are bit-identical.
//
// The element function is shared with a host build (YB_HOST_EMU; tests/host_emu/wapply_emu.cpp).  A/B option "wapply_onepass".
#pragma once

#ifdef YB_HOST_EMU
#ifndef YB_HOST_CPLX
#define YB_HOST_CPLX
struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
typedef double2 cplx;
#include <cmath>
static inline cplx cfma(cplx a, cplx b, cplx c) {
    return make_double2(std::fma(a.x, b.x, std::fma(-a.y, b.y, c.x)), std::fma(a.x, b.y, std::fma(a.y, b.x, c.y)));
}
#endif
#define YB_WAPPLY_FN static inline
#else
#define YB_WAPPLY_FN __device__ __forceinline__
#endif

namespace wapply {

constexpr int MAX_NX = 4, MAX_NY = 5;      // two-site physical index (d^2 = 4), MPO bonds up to 5 (Heisenberg)

// pos enumerates (l, r) row-major.  w: [nx][nu][ny][nv] (shared memory on the device), src: packed [nu][l][nv][r], dst: packed
// [nx][l][ny][r].
YB_WAPPLY_FN void element(long long pos, int nx, int nu, int ny, int nv, int l, int r, const cplx* w, const cplx* src, cplx* dst) {
    const int li = (int)(pos / r), ri = (int)(pos - (long long)li * r);
    cplx acc[MAX_NX][MAX_NY];
#pragma unroll
    for (int x = 0; x < MAX_NX; ++x)
#pragma unroll
        for (int y = 0; y < MAX_NY; ++y) acc[x][y] = make_double2(0.0, 0.0);
    for (int u = 0; u < nu; ++u)
        for (int v = 0; v < nv; ++v) {
            const cplx t = src[(((long long)u * l + li) * nv + v) * r + ri];
#pragma unroll
            for (int x = 0; x < MAX_NX; ++x) {
                if (x >= nx) break;
#pragma unroll
                for (int y = 0; y < MAX_NY; ++y) {
                    if (y >= ny) break;
                    const cplx wv = w[((x * nu + u) * ny + y) * nv + v];
                    if (wv.x == 0.0 && wv.y == 0.0) continue;
                    acc[x][y] = cfma(wv, t, acc[x][y]);
                }
            }
        }
#pragma unroll
    for (int x = 0; x < MAX_NX; ++x) {
        if (x >= nx) break;
#pragma unroll
        for (int y = 0; y < MAX_NY; ++y) {
            if (y >= ny) break;
            dst[(((long long)x * l + li) * ny + y) * r + ri] = acc[x][y];
        }
    }
}

}  // namespace wapply
