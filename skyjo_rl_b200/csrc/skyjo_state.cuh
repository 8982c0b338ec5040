// skyjo_state.cuh -- HBM layout of the batched SkyJo state and its bit-level helpers.
//
// Structure-of-arrays int8 layout.  Per env (N players) the state is 1 + N "planes" of 16 bytes;
// plane p of env e lives at planes[p * Bpad + e], so a warp's access to one plane is 512
// contiguous bytes (LDG.128 / STG.128, fully coalesced).  A second set ("next") holds the
// pre-dealt first state of each env's next episode (auto-reset).
//
//   plane 0      : { hdr.lo, hdr.hi, hist.lo, hist.hi }
//   plane 1 + q  : row of player q = { W0, W1, W2, W3 }
//
// row (16 bytes) -- players_cards + players_masked of one player (reference skyjo.py:63-65,
// 99-103) as int8.  Bytes 3..14 are the 12 card slots: the TRUE card value (int8), even when the
// slot is hidden; -14 where the column was removed (skyjo.py:459).  The observation header is
// 19 bytes, so in the env's observation row the cards of player q start at byte
// 19 + 12 q = 3 (mod 4): with the slots at bytes 3..14 of the 16-byte row, W1 and W2 are whole
// words of the observation and W0 / W3 each share one word with their neighbour.  Encoding an
// observation is a masked copy (hidden slots -> 15), not a per-slot transform, and revealing a
// card needs no memory access.
//   W0 bits  0..11  hidden bit per slot      (players_masked == 2)
//      bits 12..15  removed flag per column  (players_masked == 0)
//      bits 16..23  sum of the open cards + 24   (known_player_sum, skyjo.py:241-246)
//      bits 24..31  slot 0
//   W1 = slots 1..4, W2 = slots 5..8, W3 = slots 9..11, top byte 0.
//
// hist (uint64) -- the 15-bin count vector of _jit_observe_global_game_stats
// (skyjo.py:236-248), maintained incrementally: discard pile (+ open table cards in direct
// mode).  Bin j (value j-2, "code" j): 4 bits at pos(j), except the value-0 bin (j == 2), which
// also receives three zeros per column removal (skyjo.py:454-458) and is 8 bits wide:
//   pos(j) = 4j (j<2) | 8 (j==2, 8 bits) | 4j+4 (j>2)
//
// hdr (uint64):
//   0..15 step in episode (saturating)   16..19 current player   20..23 starter
//   24 phase (0 draw, 1 place)  25 terminated (frozen)  26 deck slot  27 lazy draw pile
//   28 rewards dirty  29..31 + 63 code of the draw pile's top card, prefetched (so that drawing
//   never waits for memory)  32..35 episode index mod 16  36..39 hand code (value+2, 15 none)
//   40..43 discard-top code (value+3, 0 = empty)  44..47 second-from-top code
//   48..55 cards left in the draw pile  56..62 in-game reshuffles this episode (mod 128)
//
// deck: uint8 [2][Bpad][160], one 160-byte row per env and slot.  Bytes 0..149 are the dealt
// deck as codes (value+2) in deal order: slot s of player q is deck[12q+s], the draw pile is
// deck[12N .. 12N+n_draw) in python-list order (top = last).  The row is written once per
// episode by the deal kernel; the step kernel reads one byte of it when a card is drawn from
// the pile.  Bytes 152..159: after an in-game reshuffle ("lazy" mode) the uint64 histogram of
// the cards left in the draw pile.
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

#ifndef SKYJO_TILE
#define SKYJO_TILE 32
#endif
constexpr int TILE = SKYJO_TILE;  // envs per CTA of the step / observe kernels (32, 64 or 128)
constexpr int ENV_PAD = 128;      // the env dimension of every plane is padded to a multiple of this
constexpr int PILE_ROW = 160;    // bytes per deck row
constexpr int LAZY_OFF = 152;    // offset of the lazy draw-pile histogram in a deck row
constexpr int STAT_SLOTS = 256;  // replicated stat vectors (spread atomics)
constexpr int NUM_STATS = 32;

constexpr uint64_t HDR_STEP_MASK = 0xFFFFull;
constexpr int HDR_CUR_SH = 16, HDR_STARTER_SH = 20;
constexpr uint64_t HDR_PHASE = 1ull << 24, HDR_TERMINATED = 1ull << 25, HDR_SLOT = 1ull << 26,
                   HDR_LAZY = 1ull << 27, HDR_DIRTY = 1ull << 28;
constexpr int HDR_EPLO_SH = 32, HDR_HAND_SH = 36, HDR_TOP_SH = 40, HDR_SECOND_SH = 44,
              HDR_NDRAW_SH = 48, HDR_Q_SH = 56;
constexpr uint64_t HDR_Q_MASK = 0x7Full;
constexpr uint64_t HDR_PF_MASK = (7ull << 29) | (1ull << 63);
constexpr uint32_t HAND_NONE = 15;
constexpr uint32_t BYTE_HIDDEN = 15u;     // skyjo.py:33 fill_masked_unk_value
constexpr uint32_t BYTE_REMOVED = 0xF2u;  // skyjo.py:34 fill_masked_refunded_value (-14)

SKYJO_HD constexpr int num_planes(int N) { return 1 + N; }

struct alignas(16) U128 {  // host/device stand-in for uint4
    uint32_t x, y, z, w;
};
SKYJO_HD U128 ld128(const U128 *p) {
#if defined(__CUDA_ARCH__)
    const uint4 v = *reinterpret_cast<const uint4 *>(p);
    return U128{v.x, v.y, v.z, v.w};
#else
    return *p;
#endif
}
SKYJO_HD void st128(U128 *p, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
#if defined(__CUDA_ARCH__)
    *reinterpret_cast<uint4 *>(p) = make_uint4(x, y, z, w);
#else
    *p = U128{x, y, z, w};
#endif
}

SKYJO_HD uint32_t sk_popc(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__popc(x);
#else
    return (uint32_t)__builtin_popcount(x);
#endif
}
// (hi:lo) >> sh, low 32 bits; sh in 0..31
SKYJO_HD uint32_t sk_funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, sh);
#else
    return sh ? (lo >> sh) | (hi << (32u - sh)) : lo;
#endif
}
SKYJO_HD uint64_t pack64(uint32_t lo, uint32_t hi) { return (uint64_t)lo | ((uint64_t)hi << 32); }

// prefetched code of the draw pile's top card
SKYJO_HD uint32_t hdr_pf_get(uint64_t hdr) { return ((uint32_t)(hdr >> 29) & 7u) | ((uint32_t)(hdr >> 60) & 8u); }
SKYJO_HD uint64_t hdr_pf_set(uint64_t hdr, uint32_t code) {
    return (hdr & ~HDR_PF_MASK) | ((uint64_t)(code & 7u) << 29) | ((uint64_t)(code & 8u) << 60);
}

// bit position of histogram bin for card code c (= value + 2)
SKYJO_HD uint32_t hist_pos(uint32_t c) { return 4u * c + (c > 2u ? 4u : 0u); }
SKYJO_HD uint32_t hist_get(uint64_t h, uint32_t c) {
    return (uint32_t)(h >> hist_pos(c)) & (c == 2 ? 0xFFu : 0xFu);
}
SKYJO_HD uint64_t hist_one(uint32_t c) { return 1ull << hist_pos(c); }

// 4 column flags -> 12 slot bits (column c = slots 3c..3c+2, skyjo.py:447-449)
SKYJO_HD uint32_t cols_to_slots(uint32_t f) {
    uint32_t s = (f & 1u) | ((f & 2u) << 2) | ((f & 4u) << 4) | ((f & 8u) << 6);
    return s * 7u;
}
// 4 bits -> 4 bytes of 0/1
SKYJO_HD uint32_t bits01(uint32_t x) { return ((x & 0xFu) * 0x00204081u) & 0x01010101u; }
// 4 nibbles (16 bits) -> 4 bytes
SKYJO_HD uint32_t spread4(uint32_t x) {
    x = (x | (x << 8)) & 0x00FF00FFu;
    return (x | (x << 4)) & 0x0F0F0F0Fu;
}
// 4 bits -> 4 bytes of 0x00/0xFF
SKYJO_HD uint32_t bitsFF(uint32_t x) { return bits01(x) * 0xFFu; }
// per-byte (code - 2) of four codes 0..14 (no borrow between bytes)
SKYJO_HD uint32_t codes_to_values4(uint32_t w) { return ((w | 0x80808080u) - 0x02020202u) ^ 0x80808080u; }

// ---- row accessors ---------------------------------------------------------------------------
// Four named words, never an array: a dynamically indexed register array would be demoted to
// local memory by the compiler.
struct Row {
    uint32_t w0, w1, w2, w3;
};
SKYJO_HD uint32_t row_hidden(const Row &r) { return r.w0 & 0xFFFu; }
SKYJO_HD uint32_t row_flags(const Row &r) { return (r.w0 >> 12) & 0xFu; }
SKYJO_HD uint32_t row_sum24(const Row &r) { return (r.w0 >> 16) & 0xFFu; }
SKYJO_HD void row_set_meta(Row &r, uint32_t hidden, uint32_t flags, uint32_t sum24) {
    r.w0 = (r.w0 & 0xFF000000u) | hidden | (flags << 12) | (sum24 << 16);
}
// observation byte of slot s (0..11): row byte s + 3
SKYJO_HD uint32_t row_byte(const Row &r, uint32_t s) {
    const uint32_t pos = s + 3u;
    const uint32_t lo = pos < 8u ? (pos < 4u ? r.w0 : r.w1) : (pos < 12u ? r.w2 : r.w3);
    return (lo >> (8u * (pos & 3u))) & 0xFFu;
}
SKYJO_HD void row_set_byte(Row &r, uint32_t s, uint32_t b) {
    const uint32_t pos = s + 3u, sh = 8u * (pos & 3u), k = pos >> 2;
    const uint32_t clr = ~(0xFFu << sh), val = b << sh;
    r.w0 = k == 0u ? (r.w0 & clr) | val : r.w0;
    r.w1 = k == 1u ? (r.w1 & clr) | val : r.w1;
    r.w2 = k == 2u ? (r.w2 & clr) | val : r.w2;
    r.w3 = k == 3u ? (r.w3 & clr) | val : r.w3;
}
// the three observation bytes of column c (0..3) as a 24-bit value
SKYJO_HD uint32_t row_col(const Row &r, uint32_t c) {
    // column c = row bytes 3+3c .. 5+3c: c0: (W0,W1)>>24, c1: (W1,W2)>>16, c2: W2>>8, c3: W3
    const uint32_t lo = c < 2u ? (c == 0u ? r.w0 : r.w1) : (c == 2u ? r.w2 : r.w3);
    const uint32_t hi = c < 2u ? (c == 0u ? r.w1 : r.w2) : 0u;
    const uint32_t sh = c < 2u ? (c == 0u ? 24u : 16u) : (c == 2u ? 8u : 0u);
    return sk_funnel_r(lo, hi, sh) & 0xFFFFFFu;
}
// mark column c as removed: its three bytes become -14 (skyjo.py:459)
SKYJO_HD void row_remove_col(Row &r, uint32_t c) {
    if (c == 0u) {
        r.w0 = (r.w0 & 0x00FFFFFFu) | 0xF2000000u;
        r.w1 = (r.w1 & 0xFFFF0000u) | 0x0000F2F2u;
    } else if (c == 1u) {
        r.w1 = (r.w1 & 0x0000FFFFu) | 0xF2F20000u;
        r.w2 = (r.w2 & 0xFFFFFF00u) | 0x000000F2u;
    } else if (c == 2u) {
        r.w2 = (r.w2 & 0x000000FFu) | 0xF2F2F200u;
    } else {
        r.w3 = (r.w3 & 0xFF000000u) | 0x00F2F2F2u;
    }
}
// the 12 slots as three aligned words (slots 0..3, 4..7, 8..11)
SKYJO_HD void row_cards(const Row &r, uint32_t out[3]) {
    out[0] = sk_funnel_r(r.w0, r.w1, 24);
    out[1] = sk_funnel_r(r.w1, r.w2, 24);
    out[2] = sk_funnel_r(r.w2, r.w3, 24);
}

// Observation words of a row: hidden slots read 15 (skyjo.py:259-302).  Returns the four words
// with the mask applied to the card bytes; W0's meta bits and W3's spare byte are left as they are.
SKYJO_HD uint32_t flags_to_bytemask(uint32_t f4) {  // 4 bits -> 4 bytes of 0x00 / 0xFF
#if defined(__CUDA_ARCH__)
    // bit i -> bit 8i+7, then replicate each byte's sign bit (PRMT, selector msb = sign mode;
    // __byte_perm ignores that bit, hence the PTX)
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"((f4 & 0xFu) * 0x10204080u), "r"(0u), "r"(0xBA98u));
    return d;
#else
    return bitsFF(f4);
#endif
}
SKYJO_HD Row row_observed(const Row &r) {
    const uint32_t h = row_hidden(r);
    const uint32_t m1 = flags_to_bytemask(h >> 1), m2 = flags_to_bytemask(h >> 5);
    const uint32_t m30 = flags_to_bytemask((h >> 9) | (h << 3));  // bytes 0..2: slots 9..11, byte 3: slot 0
    Row o;
    o.w0 = (r.w0 & ~(m30 & 0xFF000000u)) | (m30 & 0x0F000000u);
    o.w1 = (r.w1 & ~m1) | (m1 & 0x0F0F0F0Fu);
    o.w2 = (r.w2 & ~m2) | (m2 & 0x0F0F0F0Fu);
    o.w3 = (r.w3 & ~(m30 & 0x00FFFFFFu)) | (m30 & 0x000F0F0Fu);
    return o;
}

// legal-action bits (bit a = action a) of _jit_action_mask (skyjo.py:201-224)
SKYJO_HD uint32_t legal_bits(uint32_t hidden, uint32_t flags, bool place_phase) {
    if (!place_phase) return 3u << 24;
    const uint32_t not_removed = ~cols_to_slots(flags) & 0xFFFu;
    return not_removed | (hidden << 12);
}

struct DeviceState {
    U128 *planes;          // [num_planes][Bpad]
    U128 *next_planes;     // [num_planes][Bpad]
    uint8_t *deck;         // [2][Bpad][PILE_ROW]
    uint32_t *episode;     // [Bpad] next episode index to deal
    uint8_t *needs_deal;   // [Bpad] 0, or 1 | free_slot << 1
    unsigned long long *stats;  // [STAT_SLOTS][NUM_STATS]
    uint32_t *errflag;     // sticky consistency flag
};

struct StepParams {
    DeviceState st;
    int8_t *obs;
    int8_t *mask;
    int8_t *agent;
    uint8_t *done;
    double *reward;
    double *final_score;
    const void *actions;
    int action_dtype;
    long long B, Bpad;
    unsigned long long first_env, seed, t;
    // launches replayed from a CUDA graph (skyjo_step_random, skyjo_capi.cu): the lockstep counter of the replay's
    // first step lives in device memory and `t` is the step's offset in the replay; null otherwise
    const unsigned long long *t_base;
    double score_penalty, mean_reward, reward_refunded;
    int auto_reset, max_steps;
    int bulk_ok;  // obs / mask base pointers 16-byte aligned -> TMA bulk stores
    int pdl;      // host side: launch with programmatic stream serialization
    int pf_dist;  // step kernel: L2 prefetch distance in tiles (0 = off)
    // step kernel over a sub-range of the batch (skyjo_step_host pipelines chunks against the
    // copies): first tile and number of tiles of this launch; tiles == 0 means the whole batch
    long long tile_off, tiles;
};

// Time-major outputs of the multi-step rollout kernel: slice k of each tensor receives what a
// single fused step would have published after env-step k of the launch.
struct RolloutParams {
    int8_t *obs;     // [K][B][D]
    int8_t *mask;    // [K][B][26]
    int8_t *agent;   // [K][B]
    uint8_t *done;   // [K][B]
    int K;           // env-steps in this launch
    int publish;     // 1: the last step also writes the bound [B, ...] outputs of the handle
    int bulk_ok;     // every slice 16-byte aligned -> TMA bulk stores
};

enum : uint32_t { ERR_NEXT_NOT_READY = 1, ERR_BAD_DECK = 2, ERR_BAD_FLIPS = 4, ERR_ASSIST = 8 };

}  // namespace skyjo
