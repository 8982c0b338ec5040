// skyjo_encode.cuh -- observation + action-mask encoding and the shared-memory row stager.
//
// Implements collect_observation (reference skyjo.py:148-199): _jit_observe_global_game_stats
// (:226-257), _jit_known_player_cards(_all) (:259-302) and _jit_action_mask (:201-224) on the
// packed state of skyjo_state.cuh.  A CTA owns TILE consecutive envs, so its slice of the
// row-major outputs obs[B, D] / mask[B, 26] is one contiguous span: rows are assembled in
// shared memory (funnel-shifted to their byte offset, D is odd) and the span is written with
// one TMA bulk store (cp.async.bulk.global.shared::cta) per tensor.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "skyjo_state.cuh"

namespace skyjo {

// 4 nibbles (16 bits) -> 4 bytes
__device__ __forceinline__ uint32_t spread4(uint32_t x) {
    x = (x | (x << 8)) & 0x00FF00FFu;
    return (x | (x << 4)) & 0x0F0F0F0Fu;
}
// 4 bits -> 4 bytes of 0/1
__device__ __forceinline__ uint32_t bits01(uint32_t x) { return ((x & 0xFu) * 0x00204081u) & 0x01010101u; }
// 4 bits -> 4 bytes of 0x00/0xFF
__device__ __forceinline__ uint32_t bitsFF(uint32_t x) { return bits01(x) * 0xFFu; }

// Writes a byte stream into shared memory at an arbitrary byte offset using aligned word
// stores for the interior and byte stores at the two ragged ends (neighbouring rows of
// other threads share those words).
struct Stager {
    uint32_t *wp;
    uint32_t prev, a, sh;
    bool first;
    __device__ __forceinline__ Stager(uint8_t *base, uint32_t off) {
        a = off & 3u;
        wp = reinterpret_cast<uint32_t *>(base + (off - a));
        sh = 32u - 8u * a;
        prev = 0;
        first = true;
    }
    __device__ __forceinline__ void push(uint32_t w) {  // 4 valid bytes
        uint32_t v = __funnelshift_rc(prev, w, sh);
        if (first) {
            first = false;
            if (a == 0) {
                *wp = v;
            } else {
                uint8_t *bp = reinterpret_cast<uint8_t *>(wp);
#pragma unroll
                for (uint32_t k = 1; k < 4; ++k)
                    if (k >= a) bp[k] = (uint8_t)(v >> (8 * k));
            }
        } else {
            *wp = v;
        }
        ++wp;
        prev = w;
    }
    // flush: `r` (0..3) valid low bytes of `last`, plus the bytes of prev still pending
    template <int R>
    __device__ __forceinline__ void finish(uint32_t last) {
        uint8_t *bp = reinterpret_cast<uint8_t *>(wp);
        if (!first) {
#pragma unroll
            for (uint32_t k = 0; k < 3; ++k)
                if (k < a) bp[k] = (uint8_t)(prev >> (8 * (4 - a + k)));
        }
#pragma unroll
        for (int i = 0; i < R; ++i) bp[a + i] = (uint8_t)(last >> (8 * i));
    }
};

struct RowView {
    uint32_t w[3];   // card codes (value + 2) as bytes, slots 0..11
    uint32_t hm[3];  // 0xFF where hidden
    uint32_t hidden; // 12 bits
    uint32_t flags;  // 4 refunded-column bits
};

__device__ __forceinline__ RowView view_row(uint64_t row) {
    RowView v;
    const uint32_t lo = (uint32_t)row, hi = (uint32_t)(row >> 32);
    v.w[0] = spread4(lo & 0xFFFFu);
    v.w[1] = spread4(lo >> 16);
    v.w[2] = spread4(hi & 0xFFFFu);
    v.hidden = (hi >> 16) & 0xFFFu;
    v.flags = hi >> 28;
    v.hm[0] = bitsFF(v.hidden);
    v.hm[1] = bitsFF(v.hidden >> 4);
    v.hm[2] = bitsFF(v.hidden >> 8);
    return v;
}

// sum of the open cards of a row (players_masked == 1; refunded columns hold value 0)
__device__ __forceinline__ int open_sum(const RowView &v) {
    int acc = 0;
    acc = __dp4a((unsigned)(v.w[0] & ~v.hm[0]), 0x01010101u, (unsigned)acc);
    acc = __dp4a((unsigned)(v.w[1] & ~v.hm[1]), 0x01010101u, (unsigned)acc);
    acc = __dp4a((unsigned)(v.w[2] & ~v.hm[2]), 0x01010101u, (unsigned)acc);
    return acc - 2 * (12 - __popc(v.hidden));
}

// the 12 observation bytes of a row: value if players_masked != 2 else 15; -14 where refunded
__device__ __forceinline__ void obs_cards(const RowView &v, uint32_t out[3]) {
    uint32_t rs = v.flags ? cols_to_slots(v.flags) : 0u;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        uint32_t x = ((v.w[k] | 0x80808080u) - 0x02020202u) ^ 0x80808080u;  // per-byte code - 2
        x = (x & ~v.hm[k]) | (0x0F0F0F0Fu & v.hm[k]);                       // hidden -> 15
        if (rs) {
            uint32_t rm = bitsFF(rs >> (4 * k));
            x = (x & ~rm) | (0xF2F2F2F2u & rm);                             // refunded -> -14
        }
        out[k] = x;
    }
}

// legal-action bits (bit a = action a) of _jit_action_mask for the given row and phase
__device__ __forceinline__ uint32_t legal_bits(uint32_t hidden, uint32_t flags, bool place_phase) {
    if (!place_phase) return 3u << 24;
    uint32_t not_refunded = ~cols_to_slots(flags) & 0xFFFu;
    return not_refunded | (hidden << 12);
}

// Encodes the observation of `observer` into the CTA's shared-memory tiles.
template <int N, bool IND>
__device__ __forceinline__ void encode_rows(const uint64_t (&rows)[N], uint64_t hdr, uint64_t hist,
                                            int observer, uint8_t *s_obs, uint8_t *s_mask, int tid) {
    constexpr int D = IND ? 31 : 19 + 12 * N;
    int min_sum = 1 << 20, min_hid = 1 << 20;
    uint32_t obs_hidden = 0, obs_flags = 0;
    Stager cs(s_obs, (uint32_t)tid * D + 19u);
#pragma unroll
    for (int p = 0; p < N; ++p) {
        RowView v = view_row(rows[p]);
        min_sum = min(min_sum, open_sum(v));
        min_hid = min(min_hid, __popc(v.hidden));
        if (p == observer) {
            obs_hidden = v.hidden;
            obs_flags = v.flags;
        }
        if (!IND || p == observer) {
            uint32_t c[3];
            obs_cards(v, c);
            cs.push(c[0]);
            cs.push(c[1]);
            cs.push(c[2]);
        }
    }
    cs.template finish<0>(0u);

    // header: [min open sum (<=127), min hidden count, 15 histogram bins, discard top, hand]
    const uint32_t hlo = (uint32_t)hist, hhi = (uint32_t)(hist >> 32);
    const uint32_t e0 = spread4(hlo & 0xFFFFu), e1 = spread4(hlo >> 16);
    const uint32_t e2 = spread4(hhi & 0xFFFFu), e3 = spread4(hhi >> 16);
    const uint32_t bin2 = (hlo >> 8) & 0xFFu;
    const uint32_t hand_code = (uint32_t)(hdr >> HDR_HAND_SH) & 0xFu;
    const uint32_t top_code = (uint32_t)(hdr >> HDR_TOP_SH) & 0xFu;
    const uint32_t hand_b = hand_code == HAND_NONE ? 15u : ((hand_code - 2u) & 0xFFu);
    const uint32_t top_b = (top_code - 3u) & 0xFFu;
    const uint32_t ms = (uint32_t)min(min_sum, 127) & 0xFFu;
    Stager hs(s_obs, (uint32_t)tid * D);
    hs.push(ms | ((uint32_t)min_hid << 8) | (e0 << 16));
    hs.push(bin2 | (e1 << 8));
    hs.push((e1 >> 24) | (e2 << 8));
    hs.push((e2 >> 24) | (e3 << 8));
    hs.template finish<3>((e3 >> 24) | (top_b << 8) | (hand_b << 16));

    // action mask, 26 bytes of 0/1
    const uint32_t lb = legal_bits(obs_hidden, obs_flags, (hdr & HDR_PHASE) != 0);
    Stager ms_(s_mask, (uint32_t)tid * 26u);
#pragma unroll
    for (int k = 0; k < 6; ++k) ms_.push(bits01(lb >> (4 * k)));
    ms_.template finish<2>(bits01(lb >> 24));
}

// Copies the CTA's staged tile to global memory.  Full tiles with 16-byte aligned targets go
// out as one TMA bulk store issued by thread 0; ragged tails fall back to a byte loop.
__device__ __forceinline__ void store_tile(const uint8_t *s_src, int8_t *g_dst, uint32_t nbytes, bool bulk,
                                           int tid) {
    if (bulk) {
        if (tid == 0) {
            uint32_t saddr = (uint32_t)__cvta_generic_to_shared(s_src);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(g_dst),
                         "r"(saddr), "r"(nbytes)
                         : "memory");
        }
    } else {
        for (uint32_t i = tid; i < nbytes; i += TILE) g_dst[i] = (int8_t)s_src[i];
    }
}

__device__ __forceinline__ void bulk_commit_and_wait(int tid) {
    if (tid == 0) {
        asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
    }
}

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

}  // namespace skyjo
