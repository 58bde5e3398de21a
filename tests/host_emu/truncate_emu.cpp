// Host build of yaqs_b200/csrc/yb_truncate.cuh (test infrastructure): the SAME source k_svd_finalize calls for every keep decision,
// compiled with g++ so that it can be pinned against the reference's known answers and the oracle without a GPU.
#define YB_HOST_EMU 1
#include "../../yaqs_b200/csrc/yb_truncate.cuh"

extern "C" int truncate_keep_host(const double* s, int n, int mode, double threshold, int max_bond, int min_keep) {
    return truncate_keep(s, n, mode, threshold, max_bond, min_keep);
}
