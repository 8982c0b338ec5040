// skyjo_deal.cuh -- the deal (reset path) and the SkyjoGame-shaped debug export, per env.
//
// deal_one implements SkyjoGame.reset (reference skyjo.py:52-74): _new_drawpile (:76-82) as a
// Philox-driven Fisher-Yates over ten each of -2..12, the deal of 12 cards per player (:63-65),
// the initial discard card (:68-70, :137), _reset_card_mask (:96-103, two open slots per
// player) and _reset_start_player (:105-125, first argmax of the open sums).  It writes either
// the live planes (reset) or the "next" planes that the step kernel installs when an episode
// ends (auto-reset), plus the env's 160-byte deck row.  The deck is accessed through a small
// accessor so that the kernel can keep it in shared memory as [word][thread] (bank-conflict
// free for the data-dependent swaps) while tests/hostsim uses a plain array.
#pragma once
#include <stdint.h>
#include <string.h>

#include "../../include/skyjo_b200.h"
#include "skyjo_rng.cuh"
#include "skyjo_state.cuh"

namespace skyjo {

constexpr int DECK_WORDS = 38;  // 150 cards -> 38 words (2 pad bytes)

struct DealParams {
    DeviceState st;
    long long B, Bpad;
    unsigned long long first_env, seed;
    int N;
    int indirect;
    long long e_begin;  // flagged mode: first env of the scanned range (B is its end)
    int flagged;       // 0: all envs, 1: envs with needs_deal != 0
    int target_next;   // 0: live planes (slot 0), 1: next planes
    const int8_t *decks;   // injected decks int8[B,150] or null
    const uint8_t *flips;  // injected flips uint8[B,N,2] or null
};

// Deck: uint8_t get(int i); void set(int i, uint8_t v); uint32_t word(int w)  (i < 152, w < 38);
//       get_at(ibase, off) / set_at(ibase, off, v) = get / set(ibase + off) with ibase a multiple of 4
//       (possibly negative) and off a compile-time constant
template <class Deck>
SKYJO_HD void deal_one(const DealParams &p, long long e, uint32_t slot, Deck &deck) {
    const int N = p.N;
    const uint32_t ep = p.st.episode[e];
    p.st.episode[e] = ep + 1u;
    const unsigned long long genv = p.first_env + (unsigned long long)e;

    uint32_t fl[SKYJO_MAX_PLAYERS];  // two open slots per player, packed a | b << 4
    if (p.decks == nullptr) {
        // skyjo.py:80-81: ten each of -2..12 (codes 0..14), then shuffle
        for (int i = 0; i < 152; ++i) deck.set(i, (uint8_t)(i < 150 ? i / 10 : 0));
        // Swap k (i = 149 - k) uses word k & 3 of Philox block k >> 2.  The blocks are counter-based,
        // hence independent of each other and of the swaps: they are generated DEAL_ILP at a time,
        // one group ahead of the swaps that consume them, so that the ten dependent rounds of a
        // block overlap those of its neighbours and the shared-memory round trips of the previous
        // group's swaps (one thread's 38 blocks back to back were 60 % of the kernel's latency).
        constexpr int DEAL_ILP = 4, NBLK = (SKYJO_DECK - 1 + 3) / 4, GROUPS = (NBLK + DEAL_ILP - 1) / DEAL_ILP;
        U4 cur[DEAL_ILP], nxt[DEAL_ILP];
#pragma unroll
        for (int u = 0; u < DEAL_ILP; ++u) cur[u] = rng_block(p.seed, genv, PURPOSE_DEAL, ep, (uint32_t)u);
        // Group t swaps i = ibase + 21 - c, c = 0..15, with ibase = 128 - 16 t a multiple of 16: the
        // deck accessor resolves (ibase, constant offset) with one address computation per group, so
        // only j is addressed dynamically; the loop stays rolled (unrolled, its 130 KB of straight-line
        // code run once per warp were instruction-fetch bound).
#pragma unroll 1
        for (int t = 0; t < GROUPS; ++t) {
            const int ibase = SKYJO_DECK - 1 - 21 - 4 * DEAL_ILP * t;
#pragma unroll
            for (int u = 0; u < DEAL_ILP; ++u)
                nxt[u] = rng_block(p.seed, genv, PURPOSE_DEAL, ep, (uint32_t)(DEAL_ILP * (t + 1) + u));
#pragma unroll
            for (int u = 0; u < DEAL_ILP; ++u) {
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    const int off = 21 - (4 * u + w);
                    const int i = ibase + off;
                    if (i >= 1) {
                        const uint32_t r = w == 0 ? cur[u].x : w == 1 ? cur[u].y : w == 2 ? cur[u].z : cur[u].w;
                        const int j = (int)bounded(r, (uint32_t)(i + 1));
                        const uint8_t a = deck.get_at(ibase, off), b = deck.get(j);
                        deck.set_at(ibase, off, b);
                        deck.set(j, a);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < DEAL_ILP; ++u) cur[u] = nxt[u];
        }
        for (int q = 0; q < N; ++q) {  // skyjo.py:101 choice(12, 2, replace=False)
            U4 r = rng_block(p.seed, genv, PURPOSE_FLIPS, ep, (uint32_t)q);
            uint32_t a = bounded(r.x, 12u), b = bounded(r.y, 11u);
            if (b >= a) b += 1u;
            fl[q] = a | (b << 4);
        }
    } else {
        uint32_t cnt[15];
        for (int c = 0; c < 15; ++c) cnt[c] = 0;
        bool bad = false;
        for (int i = 0; i < SKYJO_DECK; ++i) {
            const int v = p.decks[e * SKYJO_DECK + i];
            if (v < -2 || v > 12) bad = true;
            const uint32_t c = (uint32_t)(v + 2) & 15u;
            deck.set(i, (uint8_t)c);
            for (int k = 0; k < 15; ++k) cnt[k] += (c == (uint32_t)k);
        }
        deck.set(150, 0);
        deck.set(151, 0);
        for (int c = 0; c < 15; ++c) bad |= cnt[c] > 15u;
        if (bad) {
#if defined(__CUDA_ARCH__)
            atomicOr(p.st.errflag, ERR_BAD_DECK);
#else
            *p.st.errflag |= ERR_BAD_DECK;
#endif
        }
        for (int q = 0; q < N; ++q) {
            const uint32_t a = p.flips[(e * N + q) * 2], b = p.flips[(e * N + q) * 2 + 1];
            if (a >= 12u || b >= 12u || a == b) {
#if defined(__CUDA_ARCH__)
                atomicOr(p.st.errflag, ERR_BAD_FLIPS);
#else
                *p.st.errflag |= ERR_BAD_FLIPS;
#endif
            }
            fl[q] = (a & 15u) | ((b & 15u) << 4);
        }
    }

    // rows (two open slots each, the rest hidden), open sums, histogram
    const uint32_t top_card = deck.get(SKYJO_DECK - 1);  // skyjo.py:137 pop()
    uint64_t hist = hist_one(top_card);
    U128 *dst = p.target_next ? p.st.next_planes : p.st.planes;
    int best = -1000, starter = 0;
    for (int q = 0; q < N; ++q) {
        const uint32_t a = fl[q] & 15u, b = fl[q] >> 4;
        const uint32_t ca = deck.get(12 * q + (int)a), cb = deck.get(12 * q + (int)b);
        const int sum = (int)ca + (int)cb - 4;
        if (sum > best) {  // first maximum wins (skyjo.py:112 argmax)
            best = sum;
            starter = q;
        }
        if (!p.indirect) hist += hist_one(ca) + hist_one(cb);
        // true values of the 12 slots: deck words 3q..3q+2 hold their codes
        const uint32_t v0 = codes_to_values4(deck.word(3 * q)), v1 = codes_to_values4(deck.word(3 * q + 1)),
                       v2 = codes_to_values4(deck.word(3 * q + 2));
        Row r;
        r.w0 = v0 << 24;
        r.w1 = (v0 >> 8) | (v1 << 24);
        r.w2 = (v1 >> 8) | (v2 << 24);
        r.w3 = v2 >> 8;
        row_set_meta(r, 0xFFFu & ~((1u << a) | (1u << b)), 0u, (uint32_t)(sum + 24));
        st128(dst + (long long)(1 + q) * p.Bpad + e, r.w0, r.w1, r.w2, r.w3);
    }
    const uint32_t n_draw = (uint32_t)(SKYJO_DECK - 1 - 12 * N);
    uint64_t hdr = ((uint64_t)starter << HDR_CUR_SH) | ((uint64_t)starter << HDR_STARTER_SH) |
                   (slot ? HDR_SLOT : 0ull) | (p.target_next ? HDR_DIRTY : 0ull) |
                   ((uint64_t)(ep & 15u) << HDR_EPLO_SH) | ((uint64_t)HAND_NONE << HDR_HAND_SH) |
                   ((uint64_t)(top_card + 1u) << HDR_TOP_SH) | ((uint64_t)n_draw << HDR_NDRAW_SH);
    hdr = hdr_pf_set(hdr, n_draw ? deck.get(SKYJO_DECK - 2) : 0u);  // top of the draw pile
    st128(dst + e, (uint32_t)hdr, (uint32_t)(hdr >> 32), (uint32_t)hist, (uint32_t)(hist >> 32));

    // deck row: the whole deck in deal order (38 words), lazy-pile histogram slot zeroed
    U128 *drow = reinterpret_cast<U128 *>(p.st.deck + ((long long)slot * p.Bpad + e) * PILE_ROW);
    for (int v4 = 0; v4 < PILE_ROW / 16; ++v4) {
        uint32_t w[4];
        for (int k = 0; k < 4; ++k) {
            const int src = 4 * v4 + k;
            w[k] = src < DECK_WORDS ? deck.word(src) : 0u;
        }
        st128(drow + v4, w[0], w[1], w[2], w[3]);
    }
}

// SkyjoGame-shaped dump of env e (debugging / parity tests only)
SKYJO_HD void export_one(const DeviceState &st, long long Bpad, int N, int indirect, long long e, SkyjoEnvDebug &d) {
    memset(&d, 0, sizeof(d));
    const U128 P0 = ld128(st.planes + e);
    const uint64_t hdr = pack64(P0.x, P0.y);
    uint64_t dh = pack64(P0.z, P0.w);
    const uint32_t step = (uint32_t)(hdr & HDR_STEP_MASK);
    const uint32_t starter = (uint32_t)(hdr >> HDR_STARTER_SH) & 15u;
    const uint32_t slot = (hdr & HDR_SLOT) ? 1u : 0u;
    const uint8_t *deck = st.deck + ((long long)slot * Bpad + e) * PILE_ROW;
    for (int q = 0; q < N; ++q) {
        const U128 T = ld128(st.planes + (long long)(1 + q) * Bpad + e);
        Row r;
        r.w0 = T.x;
        r.w1 = T.y;
        r.w2 = T.z;
        r.w3 = T.w;
        const uint32_t hidden = row_hidden(r), flags = row_flags(r);
        const uint32_t removed = cols_to_slots(flags);
        for (uint32_t s = 0; s < 12; ++s) {
            const bool hid = (hidden >> s) & 1u, rem = (removed >> s) & 1u;
            const int vis = (int)(int8_t)row_byte(r, s);
            d.players_cards[q][s] = (int8_t)vis;  // true value; -14 where removed
            d.players_masked[q][s] = rem ? 0 : (hid ? 2 : 1);
            if (!indirect && !hid && !rem) dh -= hist_one((uint32_t)(vis + 2));
        }
        d.num_refunded[q] = (int8_t)sk_popc(flags);
        // players rotate strictly (skyjo.py:114-120): seat q has placed once per completed turn
        const uint32_t places = step >> 1;
        const uint32_t off = ((uint32_t)q + (uint32_t)N - starter) % (uint32_t)N;
        d.num_placed[q] = (int16_t)(places / (uint32_t)N + (off < places % (uint32_t)N ? 1u : 0u));
    }
    int nd = 0;
    for (uint32_t c = 0; c < 15; ++c) {
        d.discard_hist[c] = (int8_t)hist_get(dh, c);
        nd += (int)hist_get(dh, c);
    }
    d.n_discard = (int16_t)nd;
    const uint32_t n_draw = (uint32_t)(hdr >> HDR_NDRAW_SH) & 0xFFu;
    d.n_draw = (int16_t)n_draw;
    if (hdr & HDR_LAZY) {
        d.draw_is_multiset = 1;
        const uint64_t left = *reinterpret_cast<const uint64_t *>(deck + LAZY_OFF);
        for (uint32_t c = 0; c < 15; ++c) d.draw_hist[c] = (int8_t)hist_get(left, c);
    } else {
        for (uint32_t k = 0; k < n_draw && k < SKYJO_DECK; ++k) d.drawpile[k] = (int8_t)((int)deck[12 * N + (int)k] - 2);
    }
    const uint32_t hand = (uint32_t)(hdr >> HDR_HAND_SH) & 15u, top = (uint32_t)(hdr >> HDR_TOP_SH) & 15u;
    d.hand_card = hand == HAND_NONE ? 15 : (int8_t)((int)hand - 2);
    d.discard_top = (int8_t)((int)top - 3);
    d.expected_player = (int8_t)((hdr >> HDR_CUR_SH) & 15u);
    d.expected_phase = (hdr & HDR_PHASE) ? 1 : 0;
    d.starter = (int8_t)starter;
    d.is_terminated = (hdr & HDR_TERMINATED) ? 1 : 0;
    d.n_reshuffles = (int8_t)((hdr >> HDR_Q_SH) & HDR_Q_MASK);
    d.step_in_episode = (int32_t)step;
    uint32_t ep = st.episode[e] - 1u;
    if ((ep & 15u) != ((uint32_t)(hdr >> HDR_EPLO_SH) & 15u)) ep -= 1u;
    d.episode = ep;
}

#if defined(__CUDACC__)
constexpr int DEAL_THREADS = 128;
constexpr int DEAL_SCAN = 1024;   // envs scanned per CTA in flagged mode

struct SmemDeck {  // [word][thread] layout
    uint8_t *base;
    int tid;
    __device__ __forceinline__ uint8_t get(int i) const { return base[((i >> 2) * DEAL_THREADS + tid) * 4 + (i & 3)]; }
    __device__ __forceinline__ void set(int i, uint8_t v) { base[((i >> 2) * DEAL_THREADS + tid) * 4 + (i & 3)] = v; }
    __device__ __forceinline__ uint32_t word(int w) const {
        return reinterpret_cast<const uint32_t *>(base)[w * DEAL_THREADS + tid];
    }
    __device__ __forceinline__ uint8_t get_at(int ibase, int off) const {
        return base[((ibase >> 2) * DEAL_THREADS + tid) * 4 + ((off >> 2) * DEAL_THREADS * 4 + (off & 3))];
    }
    __device__ __forceinline__ void set_at(int ibase, int off, uint8_t v) {
        base[((ibase >> 2) * DEAL_THREADS + tid) * 4 + ((off >> 2) * DEAL_THREADS * 4 + (off & 3))] = v;
    }
};

// Selection: ALL envs, or only those whose needs_deal flag is set; in the latter case a CTA
// scans DEAL_SCAN flags, compacts the hits into shared memory and deals them with dense warps.
__global__ void __launch_bounds__(DEAL_THREADS) deal_kernel(const DealParams p) {
    __shared__ __align__(16) uint8_t s_deck[DECK_WORDS * DEAL_THREADS * 4];
    __shared__ uint32_t s_list[DEAL_SCAN];
    __shared__ uint32_t s_count;
    const int tid = threadIdx.x;
    SmemDeck deck{s_deck, tid};
    if (!p.flagged) {
        const long long e = (long long)blockIdx.x * DEAL_THREADS + tid;
        if (e < p.B) deal_one(p, e, p.target_next ? 1u : 0u, deck);
        return;
    }
    const long long base = p.e_begin + (long long)blockIdx.x * DEAL_SCAN;
    if (tid == 0) s_count = 0;
    __syncthreads();
    for (int i = tid; i < DEAL_SCAN; i += DEAL_THREADS) {
        const long long e = base + i;
        if (e < p.B) {
            const uint32_t f = p.st.needs_deal[e];
            if (f) s_list[atomicAdd(&s_count, 1u)] = (uint32_t)i | ((f >> 1) << 31);
        }
    }
    __syncthreads();
    const uint32_t n = s_count;
    for (uint32_t i = tid; i < n; i += DEAL_THREADS) {
        const uint32_t item = s_list[i];
        const long long e = base + (item & 0x7FFFFFFFu);
        deal_one(p, e, item >> 31, deck);
        p.st.needs_deal[e] = 0;
    }
}

__global__ void export_kernel(const DeviceState st, long long Bpad, int N, int indirect, long long env0, long long count,
                              SkyjoEnvDebug *out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    SkyjoEnvDebug d;
    export_one(st, Bpad, N, indirect, env0 + i, d);
    out[i] = d;
}

__global__ void stats_reduce_kernel(const unsigned long long *stats, long long *out) {
    const int k = threadIdx.x;
    if (k >= NUM_STATS) return;
    unsigned long long s = 0;
    for (int slot = 0; slot < STAT_SLOTS; ++slot) s += stats[slot * NUM_STATS + k];
    out[k] = (long long)s;
}
#endif  // __CUDACC__

}  // namespace skyjo
