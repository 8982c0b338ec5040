// skyjo_state.cuh -- HBM layout of the batched SkyJo state and its bit-level helpers.
//
// Per env (N players, NP = ceil(N/2)):
//   plane 0      : uint4 = { hdr.lo, hdr.hi, hist.lo, hist.hi }
//   plane 1 + k  : uint4 = { row[2k].lo, row[2k].hi, row[2k+1].lo, row[2k+1].hi }
// Planes are structure-of-arrays over the env index: plane p of env e lives at
// planes[p * Bpad + e], so a warp's load of one plane is 512 contiguous bytes
// (LDG.128, fully coalesced).  A second set of planes ("next") holds the pre-dealt
// first state of each env's next episode.
//
// row (uint64), one per player -- replaces players_cards + players_masked
// (reference skyjo.py:63-65, 99-103):
//   bits  0..47  12 nibbles, slot i -> code = card value + 2 (0..14); true value even
//                when hidden; a refunded column holds code 2 (value 0)
//   bits 48..59  hidden bit per slot  (players_masked == 2)
//   bits 60..63  refunded flag per column (players_masked == 0, card shown as -14)
//
// hist (uint64) -- the 15-bin count vector of _jit_observe_global_game_stats
// (skyjo.py:236-248), maintained incrementally: discard pile (+ open table cards in
// direct mode).  Bin j (value j-2): 4 bits at pos(j), except the value-0 bin (j == 2),
// which also receives three zeros per column removal (skyjo.py:454-458) and is 8 bits:
//   pos(j) = 4j (j<2) | 8 (j==2, 8 bits wide) | 4j+4 (j>2)
//
// hdr (uint64):
//   0..15 step in episode (saturating)   16..19 current player   20..23 starter
//   24 phase (0 draw, 1 place)  25 terminated (frozen)  26 pile slot  27 lazy draw pile
//   28 rewards dirty  32..35 episode index mod 16  36..39 hand code (value+2, 15 none)
//   40..43 discard-top code (value+3, 0 = empty)  44..47 second-from-top code
//   48..55 cards left in the draw pile  56..63 in-game reshuffles this episode
//
// pile: uint8 [2][Bpad][160], one 160-byte row per env and slot.  Explicit mode: the draw
// pile as python-list order of codes (value+2), top = row[n_draw-1].  Lazy mode (after an
// in-game reshuffle): row[0..7] is the uint64 histogram of the cards left in the pile.
#pragma once
#include <stdint.h>

namespace skyjo {

constexpr int TILE = 128;        // envs per CTA
constexpr int PILE_ROW = 160;    // bytes per pile row
constexpr int STAT_SLOTS = 256;  // replicated stat vectors (spread atomics)
constexpr int NUM_STATS = 32;

constexpr uint64_t HDR_STEP_MASK = 0xFFFFull;
constexpr int HDR_CUR_SH = 16, HDR_STARTER_SH = 20;
constexpr uint64_t HDR_PHASE = 1ull << 24, HDR_TERMINATED = 1ull << 25, HDR_SLOT = 1ull << 26,
                   HDR_LAZY = 1ull << 27, HDR_DIRTY = 1ull << 28;
constexpr int HDR_EPLO_SH = 32, HDR_HAND_SH = 36, HDR_TOP_SH = 40, HDR_SECOND_SH = 44,
              HDR_NDRAW_SH = 48, HDR_Q_SH = 56;
constexpr uint32_t HAND_NONE = 15;

__host__ __device__ __forceinline__ constexpr int num_planes(int N) { return 1 + (N + 1) / 2; }

// bit position of histogram bin for card code c (= value + 2)
__host__ __device__ __forceinline__ uint32_t hist_pos(uint32_t c) {
    return c < 2 ? 4 * c : (c == 2 ? 8u : 4 * c + 4);
}
__host__ __device__ __forceinline__ uint32_t hist_get(uint64_t h, uint32_t c) {
    return (uint32_t)(h >> hist_pos(c)) & (c == 2 ? 0xFFu : 0xFu);
}
__host__ __device__ __forceinline__ uint64_t hist_one(uint32_t c) { return 1ull << hist_pos(c); }

__host__ __device__ __forceinline__ uint32_t row_code(uint64_t row, uint32_t slot) {
    return (uint32_t)(row >> (4 * slot)) & 0xFu;
}
__host__ __device__ __forceinline__ uint64_t row_set_code(uint64_t row, uint32_t slot, uint32_t c) {
    return (row & ~(0xFull << (4 * slot))) | ((uint64_t)c << (4 * slot));
}
// 4 column flags -> 12 slot bits (column c = slots 3c..3c+2, skyjo.py:447-449)
__host__ __device__ __forceinline__ uint32_t cols_to_slots(uint32_t f) {
    uint32_t s = (f & 1u) | ((f & 2u) << 2) | ((f & 4u) << 4) | ((f & 8u) << 6);
    return s * 7u;
}

struct DeviceState {
    uint4 *planes;         // [num_planes][Bpad]
    uint4 *next_planes;    // [num_planes][Bpad]
    uint8_t *pile;         // [2][Bpad][PILE_ROW]
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
    double score_penalty, mean_reward, reward_refunded;
    int auto_reset, max_steps;
    int bulk_ok;  // obs / mask base pointers 16-byte aligned -> TMA bulk stores
};

enum : uint32_t { ERR_NEXT_NOT_READY = 1, ERR_BAD_DECK = 2, ERR_BAD_FLIPS = 4 };

}  // namespace skyjo
