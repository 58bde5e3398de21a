// Host build of yaqs_b200/csrc/yb_philox.cuh (test infrastructure): the SAME source the kernel k_philox_uniform uses, compiled
// with g++ so that the generator can be pinned against the Random123 known-answer vectors without a GPU.
#define YB_HOST_EMU 1
#include "../../yaqs_b200/csrc/yb_philox.cuh"

extern "C" void philox_block_host(uint32_t* ctr, uint32_t k0, uint32_t k1) { philox::block(ctr, k0, k1); }

extern "C" void philox_table_host(uint64_t seed, int64_t traj, uint32_t tag, int nblocks, uint32_t block0, double* out) {
    for (int c = 0; c < nblocks; ++c) philox::pair(seed, traj, tag, block0 + (uint32_t)c, out[2 * c], out[2 * c + 1]);
}
