// Host build of yaqs_b200/csrc/yb_wapply.cuh (test infrastructure): the element function the kernel k_wapply_onepass runs per thread,
// looped over the (l, r) positions of one trajectory, so its index arithmetic is pinned against numpy without a GPU.
#define YB_HOST_EMU 1
#include "../../yaqs_b200/csrc/yb_wapply.cuh"

extern "C" void wapply_onepass_host(int nx, int nu, int ny, int nv, int l, int r, const double* w, const double* src, double* dst) {
    for (long long pos = 0; pos < (long long)l * r; ++pos)
        wapply::element(pos, nx, nu, ny, nv, l, r, (const cplx*)w, (const cplx*)src, (cplx*)dst);
}
