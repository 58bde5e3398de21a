// Host build of yaqs_b200/csrc/yb_mpo_apply.cuh (test infrastructure): the SAME element function the kernel k_mpo_site_apply
// runs per thread, looped over the output elements of one trajectory, so its index arithmetic is pinned against numpy without a GPU.
#define YB_HOST_EMU 1
#include "../../yaqs_b200/csrc/yb_mpo_apply.cuh"

extern "C" void mpo_site_apply_host(int dphys, int dl, int dr, long long l, long long r, const double* w, const double* a, double* o) {
    const long long n = l * dl * r * dr;
    for (long long e = 0; e < n; ++e)
        mpo_apply::element(e, dphys, dl, dr, l, r, (const cplx*)w, (const cplx*)a, (cplx*)o);
}
