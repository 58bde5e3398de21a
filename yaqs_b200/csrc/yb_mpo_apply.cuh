// One output element of the MPO-site x MPS-site product with fused bonds (MPO.multiply on an MPS, mpo.py:1511-1537,
// mpo_utils.py:27-56):   A'[s, (l a), (r b)] = sum_t W[s, t, a, b] A[t, l, r],   MPS bond index major in both fused bonds.
//
// The element function is shared by the kernel k_mpo_site_apply (yb_ops.cu) and a host build (YB_HOST_EMU;
// tests/host_emu/mpo_apply_emu.cpp) that pins the index arithmetic against numpy without a GPU.
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
#define YB_MPO_FN static inline
#else
#define YB_MPO_FN __device__ __forceinline__
#endif

namespace mpo_apply {

// e enumerates the (l*dl) x (r*dr) output matrix row-major; a: packed (dphys, l, r) source, o: packed (dphys, l*dl, r*dr) result,
// w: (dphys, dphys, dl, dr) MPO site tensor.  dphys <= 4.
YB_MPO_FN void element(long long e, int dphys, int dl, int dr, long long l, long long r, const cplx* w, const cplx* a, cplx* o) {
    const long long lr = l * r, ro_n = r * dr, out_lr = l * dl * ro_n;
    const long long lo = e / ro_n, ro = e - lo * ro_n;
    const long long li = lo / dl, ri = ro / dr;
    const int ia = (int)(lo - li * dl), ib = (int)(ro - ri * dr);
    cplx v[4];
    for (int t = 0; t < dphys; ++t) v[t] = a[(long long)t * lr + li * r + ri];
    for (int s = 0; s < dphys; ++s) {
        cplx acc = make_double2(0.0, 0.0);
        for (int t = 0; t < dphys; ++t) acc = cfma(w[((s * dphys + t) * dl + ia) * dr + ib], v[t], acc);
        o[(long long)s * out_lr + e] = acc;
    }
}

}  // namespace mpo_apply
