// skyjo_rng.cuh -- counter-based RNG protocol of the batched env (host + device).
//
// Replaces numba's global MT19937 (reference skyjo.py:81,94,101,135) by Philox4x32-10
// (Salmon et al., SC11) keyed by the user seed and counted by
//   c0 = global_env[31:0]
//   c1 = global_env[55:32] | purpose << 24
//   c2, c3 = purpose-specific (episode, block) / (lo, hi of t >> 2 for the policy, word t & 3)
// so that an env's games depend only on (seed, global env id), never on batch size,
// sharding or launch schedule.  Bounded integers use the multiply-shift map
// floor(r * n / 2^32) (bias <= n / 2^32 < 4e-8 for n <= 150).
#pragma once
#include <stdint.h>

#ifndef SKYJO_HD
#if defined(__CUDACC__)
#define SKYJO_HD __host__ __device__ __forceinline__
#else
#define SKYJO_HD inline
#endif
#endif

namespace skyjo {

enum : uint32_t { PURPOSE_DEAL = 1, PURPOSE_FLIPS = 2, PURPOSE_RESHUFFLE = 3, PURPOSE_POLICY = 4 };

struct U4 {
    uint32_t x, y, z, w;
};

SKYJO_HD void mulhilo(uint32_t a, uint32_t b, uint32_t &hi, uint32_t &lo) {
#if defined(__CUDA_ARCH__)
    lo = a * b;
    hi = __umulhi(a, b);
#else
    uint64_t p = (uint64_t)a * b;
    lo = (uint32_t)p;
    hi = (uint32_t)(p >> 32);
#endif
}

SKYJO_HD U4 philox4x32_10(U4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0, lo0, hi1, lo1;
        mulhilo(0xD2511F53u, c.x, hi0, lo0);
        mulhilo(0xCD9E8D57u, c.z, hi1, lo1);
        U4 n;
        n.x = hi1 ^ c.y ^ k0;
        n.y = lo1;
        n.z = hi0 ^ c.w ^ k1;
        n.w = lo0;
        c = n;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return c;
}

SKYJO_HD U4 rng_block(uint64_t seed, uint64_t env, uint32_t purpose, uint32_t a, uint32_t b) {
    U4 c;
    c.x = (uint32_t)env;
    c.y = ((uint32_t)(env >> 32) & 0xFFFFFFu) | (purpose << 24);
    c.z = a;
    c.w = b;
    return philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
}

SKYJO_HD uint32_t bounded(uint32_t r, uint32_t n) {
#if defined(__CUDA_ARCH__)
    return __umulhi(r, n);
#else
    return (uint32_t)(((uint64_t)r * n) >> 32);
#endif
}

// index (0-based, ascending) of the k-th set bit of m; k < popcount(m).  Branch-free binary
// search over popcounts (a data-dependent loop would diverge across the warp).
SKYJO_HD int nth_set_bit(uint32_t m, int k) {
    uint32_t pos = 0, kk = (uint32_t)k;
#if defined(__CUDA_ARCH__)
#define SKYJO_POPC(x) ((uint32_t)__popc(x))
#else
#define SKYJO_POPC(x) ((uint32_t)__builtin_popcount(x))
#endif
#pragma unroll
    for (uint32_t w = 16; w >= 1; w >>= 1) {
        const uint32_t c = SKYJO_POPC(m & ((1u << w) - 1u));
        const bool up = kk >= c;
        kk = up ? kk - c : kk;
        pos = up ? pos + w : pos;
        m = up ? m >> w : m;
    }
#undef SKYJO_POPC
    return (int)pos;
}

// random_admissible_policy.py:26-28: uniform over the legal actions.  The random word depends
// only on (seed, env, t), so a kernel can draw it before the state has arrived.
// One Philox block serves four consecutive lockstep steps: step t uses word t & 3 of the block counted
// by t >> 2, so the multi-step rollout kernel draws one block per four env-steps (the ten rounds were 7 %
// of its instructions).
SKYJO_HD U4 policy_block(uint64_t seed, uint64_t env, uint64_t t) {
    const uint64_t g = t >> 2;
    return rng_block(seed, env, PURPOSE_POLICY, (uint32_t)g, (uint32_t)(g >> 32));
}
SKYJO_HD uint32_t policy_word(const U4 &b, uint64_t t) {
    const uint32_t w = (uint32_t)t & 3u;
    return w == 0u ? b.x : (w == 1u ? b.y : (w == 2u ? b.z : b.w));
}
SKYJO_HD uint32_t policy_random(uint64_t seed, uint64_t env, uint64_t t) {
    return policy_word(policy_block(seed, env, t), t);
}
SKYJO_HD int policy_select(uint32_t rnd, uint32_t legal_bits) {
#if defined(__CUDA_ARCH__)
    int cnt = __popc(legal_bits);
#else
    int cnt = __builtin_popcount(legal_bits);
#endif
    return nth_set_bit(legal_bits, (int)bounded(rnd, (uint32_t)cnt));
}
SKYJO_HD int policy_pick(uint64_t seed, uint64_t env, uint64_t t, uint32_t legal_bits) {
    return policy_select(policy_random(seed, env, t), legal_bits);
}

}  // namespace skyjo
