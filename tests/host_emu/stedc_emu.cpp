// Host emulation of yaqs_b200/csrc/yb_stedc.cuh (test infrastructure): the kernel body compiled for the CPU with the threads
// of every barrier-separated phase executed one after the other.  Checks the NUMERICS of the divide & conquer solver against
// LAPACK without a GPU; the GPU tests check the parallel execution of the same source.
#define YB_HOST_EMU 1
#include <cstring>
#include <memory>
using std::min;
using std::max;
#include "../../yaqs_b200/csrc/yb_stedc.cuh"

extern "C" int stedc_host(double* d, const double* e, int k, double* zt_out, int nthr) {
    if (k > 512 || k < 1) return -1;
    std::unique_ptr<stedc::Smem<512>> S(new stedc::Smem<512>());
    std::memset(S.get(), 0, sizeof(stedc::Smem<512>));
    std::vector<double> qa((size_t)k * k, -7.0), qb((size_t)k * k, -9.0), acc((size_t)nthr * 32, 0.0);
    int fail = 0;
    double* q = stedc::solve<512>(*S, d, e, k, qa.data(), qb.data(), nthr, &fail, acc);
    std::memcpy(zt_out, q, sizeof(double) * (size_t)k * k);
    return fail;
}
