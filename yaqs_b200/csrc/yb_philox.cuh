// Philox4x32-10 counter-based uniforms for the jump tests (device RNG of rng_mode = "philox").
//
// Role on the path: the reference gives every trajectory its own seeded numpy stream (core/random_utils.py:20-87:
// SeedSequence([seed, traj, TAG]) for initialize/step_through, SeedSequence([seed, traj, timestep, TAG]) per measurement
// copy) and a jump test consumes one double, plus one more if it jumps (methods/stochastic_process.py:190-292).  The
// counter-based equivalent keeps exactly that structure: key = the 64-bit run seed, counter = (block, tag, traj_lo, traj_hi);
// one Philox block = 128 bits = the two doubles of one jump test.  A stream is therefore addressed, not advanced: the
// table of a trajectory depends only on (seed, global trajectory index), never on batch composition or GPU count.
//
// The same source compiles for the host (YB_HOST_EMU; tests/host_emu/philox_emu.cpp) where it is pinned against the
// Random123 known-answer vectors and a numpy restatement.
#pragma once
#include <stdint.h>

#ifdef YB_HOST_EMU
#define YB_PHILOX_FN static inline
#else
#define YB_PHILOX_FN __host__ __device__ __forceinline__
#endif

namespace philox {

constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;      // round multipliers
constexpr uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;      // Weyl key increments (golden ratio, sqrt(3) - 1)

YB_PHILOX_FN void block(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)M0 * c[0], p1 = (uint64_t)M1 * c[2];
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        c[1] = (uint32_t)p1;
        c[3] = (uint32_t)p0;
        c[0] = n0;
        c[2] = n2;
        k0 += W0;
        k1 += W1;
    }
}

// 53-bit double in [0, 1) from two 32-bit words (the construction numpy uses for 32-bit generators)
YB_PHILOX_FN double to_double(uint32_t hi, uint32_t lo) {
    return (double)(((uint64_t)(hi >> 5) << 26) | (uint64_t)(lo >> 6)) * (1.0 / 9007199254740992.0);
}

// the two doubles of block `blk` of stream (seed, traj, tag)
YB_PHILOX_FN void pair(uint64_t seed, int64_t traj, uint32_t tag, uint32_t blk, double& u0, double& u1) {
    uint32_t c[4] = {blk, tag, (uint32_t)((uint64_t)traj & 0xFFFFFFFFu), (uint32_t)((uint64_t)traj >> 32)};
    block(c, (uint32_t)(seed & 0xFFFFFFFFu), (uint32_t)(seed >> 32));
    u0 = to_double(c[0], c[1]);
    u1 = to_double(c[2], c[3]);
}

}  // namespace philox
