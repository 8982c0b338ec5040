// skyjo_core.cuh -- the per-env body of the fused step kernel: state transition, end-of-game
// scoring, auto-reset install and the observation / action-mask word streams.
//
// Everything here is written per env ("one game per thread") against the layout of
// skyjo_state.cuh and is __host__ __device__ so that tests/hostsim can compile the same source
// with g++ and check it against the oracle without a GPU (test infrastructure only: the product
// runs these functions on the device, inside step_kernel / observe_kernel).
//
// Reference semantics implemented (file:line into /root/reference/rlskyjo):
//   SkyjoGame.act game/skyjo.py:308-335, _action_draw_card :337-374 (last-round check :350-356,
//   draw-pile reshuffle :361-365), _action_place :376-427, column removal :431-469,
//   _evaluate_game :477-498, SimpleSkyjoEnv._calc_final_rewards environment/skyjo_env.py:293-312,
//   TerminateIllegalWrapper(illegal_reward=-1) :23, collect_observation game/skyjo.py:148-199,
//   _jit_action_mask :201-224, random_admissible_policy.py:26-28.
#pragma once
#include <stdint.h>

#include "../../include/skyjo_b200.h"
#include "skyjo_rng.cuh"
#include "skyjo_state.cuh"

#ifndef SKYJO_DRAW_SHORTCUT_MAX_N
#define SKYJO_DRAW_SHORTCUT_MAX_N 4
#endif

namespace skyjo {

template <int N>
struct Env {
    uint64_t hdr, hist;
    Row row[N];
};

// what one step did, for the statistics vector
struct Outcome {
    int done_code;   // SKYJO_RUNNING / SKYJO_DONE_*
    int act_class;   // -1 none, 0: action 24, 1: action 25, 2: swap, 3: flip
    int reshuffled;  // in-game reshuffle happened
    // valid when done_code == SKYJO_DONE_GAME_OVER and the env was not frozen:
    int scored, ep_steps, raw_sum, winner_raw, fin_raw, penalised, refunds, winner, starter0;
    uint32_t dirty_rows;  // bit q: row q changed (all rows after an install)
    // Code of the draw pile's new top card when action 24 consumed the prefetched one (else
    // PF_KEEP).  It comes from a byte load issued by env_step and is merged into the header only
    // by store_env, so nothing else in the step waits for that load.
    uint32_t pf_new;
};
constexpr uint32_t PF_KEEP = 0xFFFFFFFFu;

// Results the WARP computed for this env before env_step ran (device only, skyjo_step.cuh
// warp_assist): the two rare events whose scalar code is long -- end-of-game scoring (N rows) and
// the discard-only histogram an in-game reshuffle needs in direct mode (12 N slots) -- are executed
// by one or two lanes of a warp while the other thirty wait.  The kernels hand such an env's rows
// to lanes 0..N-1, one row each, and pass the reduced results in here.
struct Assist {
    int scored;  // 1: rewards and final scores of the ending game are stored, the fields below are valid
    int raw_sum, winner_raw, fin_raw, penalised, refunds, winner;
    int has_dh;   // 1: dh = the 15-bin histogram of the discard pile alone
    uint64_t dh;
};

SKYJO_HD double sk_dadd(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
SKYJO_HD double sk_dmul(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
SKYJO_HD double sk_ddiv(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __ddiv_rn(a, b);
#else
    return a / b;
#endif
}
SKYJO_HD void sk_flag(uint32_t *flag, uint32_t bit) {
#if defined(__CUDA_ARCH__)
    atomicOr(flag, bit);
#else
    *flag |= bit;
#endif
}

// numpy's add.reduce order for a float64 vector of N <= 12 entries
template <int N>
SKYJO_HD double np_sum(const double (&a)[N]) {
    if constexpr (N < 8) {
        double r = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) r = sk_dadd(r, a[i]);
        return r;
    } else {
        double r = sk_dadd(sk_dadd(sk_dadd(a[0], a[1]), sk_dadd(a[2], a[3])),
                           sk_dadd(sk_dadd(a[4], a[5]), sk_dadd(a[6], a[7])));
#pragma unroll
        for (int i = 8; i < N; ++i) r = sk_dadd(r, a[i]);
        return r;
    }
}

// sample one card code from a packed histogram (sampling without replacement of the
// reshuffled pile, DESIGN.md "in-game reshuffle"): the smallest code c with
// cnt[0] + .. + cnt[c] > idx; idx < total <= 150.  Branch-free SWAR instead of a 14-step scan (the
// scan was 95 warp-instructions per draw-slot step at N=8, run by one or two lanes of most warps):
// the 16 nibbles are spread to bytes, byte-wise inclusive prefix sums come from multiplications by
// 0x01010101 (no carries: every prefix <= 150), and "prefix > idx" is the carry out of
// prefix + (255 - idx), assembled per byte from the 7-bit sum and the two top bits.
SKYJO_HD uint32_t hist_take(uint64_t &h, uint32_t idx) {
    const uint32_t lo = (uint32_t)h, hi = (uint32_t)(h >> 32);
    uint32_t b0 = spread4(lo & 0xFFFFu);
    const uint32_t b1 = spread4(lo >> 16), b2 = spread4(hi & 0xFFFFu), b3 = spread4(hi >> 16);
    b0 = (b0 & 0x00FFFFFFu) + ((b0 >> 24) << 20);  // bin 2 is 8 bits wide: byte 2 += 16 * byte 3, byte 3 = 0
    uint32_t p[4];
    p[0] = b0 * 0x01010101u;
    p[1] = b1 * 0x01010101u + (p[0] >> 24) * 0x01010101u;
    p[2] = b2 * 0x01010101u + (p[1] >> 24) * 0x01010101u;
    p[3] = b3 * 0x01010101u + (p[2] >> 24) * 0x01010101u;
    const uint32_t k = 255u - idx, kl = (k & 0x7Fu) * 0x01010101u, m7 = (k & 0x80u) ? 0xFFFFFFFFu : 0u;
    uint32_t n = 16u;  // index of the first byte whose prefix exceeds idx
#pragma unroll
    for (int w = 0; w < 4; ++w) {
        const uint32_t x = p[w], sm = (x & 0x7F7F7F7Fu) + kl;
        n -= sk_popc(((x & sm) | ((x | sm) & m7)) & 0x80808080u);
    }
    const uint32_t code = n - (n > 2u ? 1u : 0u);  // byte 3 is the empty upper half of bin 2
    h -= hist_one(code);
    return code;
}

SKYJO_HD uint32_t hist_total(uint64_t h) {
    const uint32_t lo = (uint32_t)h, hi = (uint32_t)(h >> 32);
    const uint32_t a = (lo & 0x0F0F0F0Fu) + ((lo >> 4) & 0x0F0F0F0Fu), b = (hi & 0x0F0F0F0Fu) + ((hi >> 4) & 0x0F0F0F0Fu);
    // nibble 3 of lo is the high half of the 8-bit bin 2: it weighs 16, the nibble sum counted it once
    return (((a + b) * 0x01010101u) >> 24) + 15u * ((lo >> 12) & 0xFu);
}

// raw (unpenalised) score of 12 true card values given as three words of int8:
// columns whose three cards are not all equal (skyjo.py:488-493)
SKYJO_HD int score12(const uint32_t v[3]) {
    int s = 0;
#pragma unroll
    for (int col = 0; col < 4; ++col) {
        int b[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int i = 3 * col + k;
            b[k] = (int)(int8_t)((v[i >> 2] >> (8 * (i & 3))) & 0xFFu);
        }
        if (!(b[0] == b[1] && b[1] == b[2])) s += b[0] + b[1] + b[2];
    }
    return s;
}

SKYJO_HD int load_action(const void *actions, int dtype, long long e) {
    switch (dtype) {
        case SKYJO_ACT_U8: return (int)reinterpret_cast<const uint8_t *>(actions)[e];
        case SKYJO_ACT_I8: return (int)reinterpret_cast<const int8_t *>(actions)[e];
        case SKYJO_ACT_I32: return reinterpret_cast<const int32_t *>(actions)[e];
        default: {
            long long a = reinterpret_cast<const long long *>(actions)[e];
            return (a < 0 || a > 255) ? 255 : (int)a;
        }
    }
}

// episode index of the running episode of env e (episode[] holds the next index to deal; one
// or two episodes may have been dealt ahead)
SKYJO_HD uint32_t running_episode(const DeviceState &st, long long e, uint64_t hdr) {
    uint32_t ep = st.episode[e] - 1u;
    if ((ep & 15u) != ((uint32_t)(hdr >> HDR_EPLO_SH) & 15u)) ep -= 1u;
    return ep;
}

template <int N>
SKYJO_HD void load_env(const U128 *planes, long long Bpad, long long e, Env<N> &s) {
    const U128 P0 = ld128(planes + e);
    s.hdr = pack64(P0.x, P0.y);
    s.hist = pack64(P0.z, P0.w);
#pragma unroll
    for (int q = 0; q < N; ++q) {
        const U128 T = ld128(planes + (long long)(1 + q) * Bpad + e);
        s.row[q].w0 = T.x;
        s.row[q].w1 = T.y;
        s.row[q].w2 = T.z;
        s.row[q].w3 = T.w;
    }
}

template <int N>
SKYJO_HD void store_env(U128 *planes, long long Bpad, long long e, const Env<N> &s, uint32_t dirty_rows,
                        uint32_t pf_new = 0xFFFFFFFFu) {
    const uint64_t hdr = pf_new == 0xFFFFFFFFu ? s.hdr : hdr_pf_set(s.hdr, pf_new);
    st128(planes + e, (uint32_t)hdr, (uint32_t)(hdr >> 32), (uint32_t)s.hist, (uint32_t)(s.hist >> 32));
#pragma unroll
    for (int q = 0; q < N; ++q)
        if ((dirty_rows >> q) & 1u)
            st128(planes + (long long)(1 + q) * Bpad + e, s.row[q].w0, s.row[q].w1, s.row[q].w2, s.row[q].w3);
}

// End-of-game scoring of one env by its own thread (skyjo.py:477-498, skyjo_env.py:293-312): raw scores, the
// finisher's penalty, float64 rewards in numpy's summation order; writes reward / final_score rows and the
// statistics fields of `oc`.  About one lane in 66 reaches it in a draw slot at N = 4, i.e. 39 % of the warps run
// it with one or two active lanes.  SKYJO_SCORE_OUTLINE = 1 keeps it out of line on the device (a call instead of
// 350 inlined instructions between the draw and the place transition); measured in DESIGN.md section 8.
#ifndef SKYJO_SCORE_OUTLINE
#define SKYJO_SCORE_OUTLINE 0
#endif
#if defined(__CUDACC__) && SKYJO_SCORE_OUTLINE
#define SKYJO_SCORE_FN __host__ __device__ __noinline__
#else
#define SKYJO_SCORE_FN SKYJO_HD
#endif
#if defined(__CUDACC__) && SKYJO_SCORE_OUTLINE == 2  // by value: the caller's state stays in registers
#define SKYJO_SCORE_ENV(N) const Env<N> s
#else
#define SKYJO_SCORE_ENV(N) const Env<N> &s
#endif
template <int N>
SKYJO_SCORE_FN void score_game(const StepParams &p, long long e, SKYJO_SCORE_ENV(N), int cur, Outcome &oc) {
    int raw[N];
    int mn = 1 << 30, refunds = 0, raw_sum = 0;
#pragma unroll
    for (int q = 0; q < N; ++q) {
        const Row &r = s.row[q];
        uint32_t v[3];
        row_cards(r, v);  // hidden cards count at their true value (skyjo.py:488-493)
        raw[q] = score12(v);
        mn = raw[q] < mn ? raw[q] : mn;
        raw_sum += raw[q];
        refunds += (int)sk_popc(row_flags(r));
    }
    int fin_raw = raw[0];
#pragma unroll
    for (int q = 1; q < N; ++q)
        if (q == cur) fin_raw = raw[q];
    const bool penalised = mn != fin_raw;  // skyjo.py:496
    double score[N];
#pragma unroll
    for (int q = 0; q < N; ++q) {
        score[q] = (double)raw[q];
        if (penalised && q == cur) score[q] = sk_dmul(score[q], p.score_penalty);
    }
    // skyjo_env.py:307-311
    const double mean = sk_ddiv(np_sum<N>(score), (double)N);
    int winner = 0;
    double best = score[0];
#pragma unroll
    for (int q = 0; q < N; ++q) {
        double r = sk_dadd(sk_dadd(-score[q], mean), p.mean_reward);
        if (p.reward_refunded != 0.0)
            r = sk_dadd(r, sk_dmul((double)sk_popc(row_flags(s.row[q])), p.reward_refunded));
        p.reward[e * N + q] = r;
        p.final_score[e * N + q] = score[q];
        if (score[q] < best) {
            best = score[q];
            winner = q;
        }
    }
    oc.raw_sum = raw_sum;
    oc.winner_raw = mn;
    oc.fin_raw = fin_raw;
    oc.penalised = penalised ? 1 : 0;
    oc.refunds = refunds;
    oc.winner = winner;
}

// One env-step for env e: SkyjoGame.act + rewards + (on episode end) auto-reset install.
// `s` holds the loaded state and is updated in place; the caller stores it back.  With POLICY
// the action is the uniform legal choice selected by `policy_rnd` = policy_random(seed, env, t).
// With ASSIST the caller guarantees `as` carries the results of every rare event this step runs
// into (a missing one raises ERR_ASSIST); without it (host build, tests/hostsim) the scalar code runs.
// With DEFER (device only, N below the assist threshold) a game that ends in the phase-locked reset mode is NOT
// scored here: oc.scored = 2 asks the caller to score it with the whole warp right after this call
// (skyjo_step.cuh warp_score_deferred) -- the rows stay in `s` because that mode installs the next episode one
// slot later.  The other reset modes replace the rows below and keep the scalar scoring.
template <int N, bool IND, bool POLICY, bool ASSIST = false, bool DEFER = false>
SKYJO_HD Outcome env_step(const StepParams &p, long long e, Env<N> &s, int action, uint32_t policy_rnd = 0u,
                          const Assist *as = nullptr) {
    Outcome oc;
    oc.done_code = SKYJO_RUNNING;
    oc.act_class = -1;
    oc.reshuffled = 0;
    oc.scored = 0;
    oc.dirty_rows = 0;
    oc.pf_new = PF_KEEP;
    uint64_t hdr = s.hdr, hist = s.hist;
    const int cur = (int)(hdr >> HDR_CUR_SH) & 0xF;
    const bool place_phase = (hdr & HDR_PHASE) != 0;

    // auto_reset == 2 ("next step", phase-locked): an episode that ended in a draw slot is not
    // replaced in the step that ended it; the env publishes its terminal observation and spends the
    // NEXT lockstep slot on the reset (action ignored, not an env-step -- the reference's separate
    // reset() call between two games, sample_game.py:8-9).  An episode lasts an odd number of
    // act() calls, so every env of the batch is in the same phase at every step and warps never
    // diverge between the draw and the place transition.
    const bool reset_slot = (hdr & HDR_TERMINATED) && p.auto_reset == 2;
    if ((hdr & HDR_DIRTY) || reset_slot) {
        // the previous step ended an episode: its rewards have been consumed
#pragma unroll
        for (int q = 0; q < N; ++q) p.reward[e * N + q] = 0.0;
        hdr &= ~HDR_DIRTY;
    }
    if ((hdr & HDR_TERMINATED) && !reset_slot) {
        oc.done_code = SKYJO_DONE_GAME_OVER;  // skyjo.py:316-321: playing a finished game returns True
        s.hdr = hdr;
        return oc;
    }

    do {  // skipped in a reset slot
        if (reset_slot) break;
        Row a = s.row[0];
#pragma unroll
        for (int q = 1; q < N; ++q)
            if (q == cur) a = s.row[q];
        uint32_t hidden = row_hidden(a), flags = row_flags(a);
        const uint32_t legal = legal_bits(hidden, flags, place_phase);
        const unsigned long long genv = p.first_env + (unsigned long long)e;
        // uniform over the legal actions; in the draw phase these are exactly {24, 25}, where the k-th set bit
        // of policy_select is 24 + k (a warp-uniform shortcut when the batch is phase-locked)
        // (small N only: from N = 5 the extra branch costs registers -- spills, +5 % kernel time at N = 8)
        if (POLICY)
            action = (N <= SKYJO_DRAW_SHORTCUT_MAX_N && !place_phase) ? 24 + (int)bounded(policy_rnd, 2u)
                                                                      : policy_select(policy_rnd, legal);
        const bool is_legal = action >= 0 && action < 26 && ((legal >> action) & 1u);
        uint32_t step = (uint32_t)(hdr & HDR_STEP_MASK);
        const uint8_t *deck = p.st.deck + (((hdr & HDR_SLOT) ? p.Bpad : 0ll) + e) * PILE_ROW;

        if (!is_legal) {
            // TerminateIllegalWrapper(illegal_reward=-1), skyjo_env.py:23
            oc.done_code = SKYJO_DONE_ILLEGAL;
#pragma unroll
            for (int q = 0; q < N; ++q) p.reward[e * N + q] = (q == cur) ? -1.0 : 0.0;
        } else if (!place_phase) {
            oc.act_class = action - 24;
            if (hidden == 0) {
                // ---- game over (skyjo.py:350-356): score, penalty, rewards -----------------------
                oc.done_code = SKYJO_DONE_GAME_OVER;
                oc.scored = 1;
                oc.ep_steps = (int)step + 1;
                if (ASSIST) {
                    // scored by the warp (warp_assist): rewards and final scores are already in memory
                    if (!as->scored) sk_flag(p.st.errflag, ERR_ASSIST);
                    oc.raw_sum = as->raw_sum;
                    oc.winner_raw = as->winner_raw;
                    oc.fin_raw = as->fin_raw;
                    oc.penalised = as->penalised;
                    oc.refunds = as->refunds;
                    oc.winner = as->winner;
                } else if (DEFER && p.auto_reset == 2) {
                    oc.scored = 2;
                } else {
                    score_game<N>(p, e, s, cur, oc);
                }
                oc.starter0 = (((hdr >> HDR_STARTER_SH) & 0xF) == 0) ? 1 : 0;
            } else {
                uint32_t code;
                if (action == 24) {
                    // ---- draw from the draw pile (skyjo.py:359-366) -------------------------------
                    uint32_t n_draw = (uint32_t)(hdr >> HDR_NDRAW_SH) & 0xFFu;
                    uint64_t *lazy = reinterpret_cast<uint64_t *>(const_cast<uint8_t *>(deck) + LAZY_OFF);
                    if (n_draw == 0) {
                        // reshuffle the whole discard pile into a new draw pile (:361-365)
                        uint64_t dh = hist;
                        if (!IND && ASSIST) {
                            if (!as->has_dh) sk_flag(p.st.errflag, ERR_ASSIST);
                            dh = as->dh;
                        } else if (!IND) {  // direct-mode hist also counts the open table cards
#pragma unroll
                            for (int q = 0; q < N; ++q) {
                                const Row &r = s.row[q];
                                const uint32_t open = ~(row_hidden(r) | cols_to_slots(row_flags(r))) & 0xFFFu;
                                for (uint32_t sl = 0; sl < 12; ++sl)
                                    if ((open >> sl) & 1u) dh -= hist_one((row_byte(r, sl) + 2u) & 0xFFu);
                            }
                        }
                        const uint32_t total = hist_total(dh);
                        const uint32_t ep = running_episode(p.st, e, hdr);
                        const uint32_t q8 = (uint32_t)(hdr >> HDR_Q_SH) & (uint32_t)HDR_Q_MASK;
                        U4 r0 = rng_block(p.seed, genv, PURPOSE_RESHUFFLE, ep, (q8 << 16) | total);
                        uint64_t left = dh;
                        const uint32_t e0 = hist_take(left, bounded(r0.x, total));
                        hist = hist - dh + hist_one(e0);  // new discard pile = [e0]
                        hdr = (hdr & ~((0xFFull << HDR_TOP_SH) | (HDR_Q_MASK << HDR_Q_SH))) |
                              ((uint64_t)(e0 + 1u) << HDR_TOP_SH) | ((uint64_t)((q8 + 1u) & (uint32_t)HDR_Q_MASK) << HDR_Q_SH) |
                              HDR_LAZY;
                        n_draw = total - 1u;
                        U4 r1 = rng_block(p.seed, genv, PURPOSE_RESHUFFLE, ep, (q8 << 16) | n_draw);
                        code = hist_take(left, bounded(r1.x, n_draw));
                        *lazy = left;
                        oc.reshuffled = 1;
                    } else if (hdr & HDR_LAZY) {
                        uint64_t left = *lazy;
                        const uint32_t ep = running_episode(p.st, e, hdr);
                        const uint32_t q8 = ((uint32_t)(hdr >> HDR_Q_SH) - 1u) & (uint32_t)HDR_Q_MASK;
                        U4 r1 = rng_block(p.seed, genv, PURPOSE_RESHUFFLE, ep, (q8 << 16) | n_draw);
                        code = hist_take(left, bounded(r1.x, n_draw));
                        *lazy = left;
                    } else {
                        // the top card was prefetched into the header; fetch the one below it for the
                        // next draw (its value is only needed when the header is stored)
                        code = hdr_pf_get(hdr);
                        oc.pf_new = n_draw > 1u ? (uint32_t)deck[12 * N + n_draw - 2u] : 0u;
                    }
                    n_draw -= 1u;
                    hdr = (hdr & ~(0xFFull << HDR_NDRAW_SH)) | ((uint64_t)n_draw << HDR_NDRAW_SH);
                } else {
                    // ---- take the discard top (skyjo.py:370) ----------------------------------------
                    const uint32_t top = (uint32_t)(hdr >> HDR_TOP_SH) & 0xFu;
                    const uint32_t second = (uint32_t)(hdr >> HDR_SECOND_SH) & 0xFu;
                    code = top - 1u;
                    hist -= hist_one(code);
                    hdr = (hdr & ~(0xFFull << HDR_TOP_SH)) | ((uint64_t)second << HDR_TOP_SH);
                }
                hdr = (hdr & ~(0xFull << HDR_HAND_SH)) | ((uint64_t)code << HDR_HAND_SH) | HDR_PHASE;
            }
        } else {
            // ---- place (skyjo.py:376-427) --------------------------------------------------------
            const uint32_t hand = (uint32_t)(hdr >> HDR_HAND_SH) & 0xFu;
            uint32_t top = (uint32_t)(hdr >> HDR_TOP_SH) & 0xFu;
            uint32_t second = top;
            const bool swap = action < 12;
            const uint32_t sl = swap ? (uint32_t)action : (uint32_t)action - 12u;
            const bool was_hidden = (hidden >> sl) & 1u;
            const uint32_t tc = (row_byte(a, sl) + 2u) & 0xFFu;  // code of the card in the slot
            int sum24 = (int)row_sum24(a);
            if (swap) {  // :389-395 the old card (even if hidden) goes to the discard pile
                oc.act_class = 2;
                top = tc + 1u;
                if (IND || was_hidden) hist += hist_one(tc);
                if (!IND) hist += hist_one(hand);
                row_set_byte(a, sl, (hand - 2u) & 0xFFu);
                sum24 += (int)hand - 2 - (was_hidden ? 0 : (int)tc - 2);
            } else {  // :396-404 discard the hand card and reveal the slot
                oc.act_class = 3;
                top = hand + 1u;
                hist += hist_one(hand);
                if (!IND) hist += hist_one(tc);
                sum24 += (int)tc - 2;
            }
            hidden &= ~(1u << sl);
            // column removal (:431-469): only the touched column can newly qualify
            const uint32_t col = sl / 3u;
            const uint32_t c3 = row_col(a, col);
            const uint32_t b0 = c3 & 0xFFu;
            if (c3 == b0 * 0x010101u && ((hidden >> (3u * col)) & 7u) == 0u && !((flags >> col) & 1u)) {
                const uint32_t c0 = (b0 + 2u) & 0xFFu;
                if (!IND) hist -= 3ull * hist_one(c0);
                hist += 3ull * hist_one(2u);  // three zeros go to the discard pile (:454-458)
                top = 3u;
                second = 3u;
                flags |= 1u << col;
                row_remove_col(a, col);
                sum24 -= 3 * ((int)c0 - 2);
            }
            row_set_meta(a, hidden, flags, (uint32_t)sum24);
            const int nxt = (cur + 1 == N) ? 0 : cur + 1;
            hdr = (hdr & ~((0xFull << HDR_CUR_SH) | HDR_PHASE | (0xFFFull << HDR_HAND_SH))) |
                  ((uint64_t)nxt << HDR_CUR_SH) | ((uint64_t)HAND_NONE << HDR_HAND_SH) |
                  ((uint64_t)top << HDR_TOP_SH) | ((uint64_t)second << HDR_SECOND_SH);
#pragma unroll
            for (int q = 0; q < N; ++q)
                if (q == cur) s.row[q] = a;
            oc.dirty_rows |= 1u << cur;
        }
        if (oc.done_code == SKYJO_RUNNING) {
            step = step + 1u < 0xFFFFu ? step + 1u : 0xFFFFu;
            hdr = (hdr & ~HDR_STEP_MASK) | step;
            if (p.max_steps > 0 && step >= (uint32_t)p.max_steps) {
                oc.done_code = SKYJO_DONE_TRUNCATED;
#pragma unroll
                for (int q = 0; q < N; ++q) p.reward[e * N + q] = 0.0;
            }
        }
    } while (0);
    // ---- episode end: install the pre-dealt next episode, defer it to the next slot, or freeze ----
    if (oc.done_code != SKYJO_RUNNING || reset_slot) {
        bool installed = false;
        // phase-locked mode: an end in a draw slot (game over, illegal draw, truncation after a draw)
        // is deferred; an end in a place slot (illegal place, truncation after a place) installs at
        // once -- either way the new episode's first draw falls on a draw slot of the batch
        const bool defer = p.auto_reset == 2 && !reset_slot && !place_phase;
        if (p.auto_reset && !defer) {
            // all 1 + N planes of the pre-dealt episode are requested at once (one memory round
            // trip, L2 hits when the previous step prefetched them); the episode tag is checked
            // on the loaded header
            Env<N> nx;
            load_env<N>(p.st.next_planes, p.Bpad, e, nx);
            const uint32_t want = ((uint32_t)(hdr >> HDR_EPLO_SH) + 1u) & 15u;
            if (((uint32_t)(nx.hdr >> HDR_EPLO_SH) & 15u) == want) {
                const uint32_t old_slot = (hdr & HDR_SLOT) ? 1u : 0u;
                s = nx;
                oc.pf_new = PF_KEEP;
                // DIRTY: the rewards of the ended episode are cleared by the next step (in a reset
                // slot they were cleared above, one step after they were published)
                hdr = reset_slot ? (s.hdr & ~HDR_DIRTY) : (s.hdr | HDR_DIRTY);
                hist = s.hist;
                oc.dirty_rows = (1u << N) - 1u;
                p.st.needs_deal[e] = (uint8_t)(1u | (old_slot << 1));
                installed = true;
            } else {
                sk_flag(p.st.errflag, ERR_NEXT_NOT_READY);
            }
        }
        if (!installed) hdr |= HDR_TERMINATED;
    }
    s.hdr = hdr;
    s.hist = hist;
    return oc;
}

// ---- observation / action-mask word streams ---------------------------------------------------
// An env's observation row is D = 19 + 12 R bytes (R = N rows in direct mode, 1 in indirect
// mode) = NW words with 3 valid bytes in the last; its mask row is 26 bytes = 7 words with 2
// valid bytes in the last.  The streams are relative to the row start; staging to the row's
// (unaligned) byte offset in the output tile is done by stage_stream (skyjo_encode.cuh).
template <int N, bool IND>
struct ObsWords {
    static constexpr int R = IND ? 1 : N;
    static constexpr int NW = 5 + 3 * R;
    static constexpr int D = 19 + 12 * R;
    uint32_t s[NW];
    uint32_t m[7];
};

// collect_observation (skyjo.py:148-199) of `observer` on state s.
template <int N, bool IND>
SKYJO_HD void encode_words(const Env<N> &s, int observer, ObsWords<N, IND> &o, uint32_t *observer_hidden = nullptr) {
    uint32_t min_sum24 = 255u, min_hid = 12u;
#pragma unroll
    for (int q = 0; q < N; ++q) {
        const uint32_t sm = row_sum24(s.row[q]), hd = sk_popc(row_hidden(s.row[q]));
        min_sum24 = sm < min_sum24 ? sm : min_sum24;
        min_hid = hd < min_hid ? hd : min_hid;
    }
    Row ob = s.row[0];
#pragma unroll
    for (int q = 1; q < N; ++q)
        if (q == observer) ob = s.row[q];

    // header: [min open sum (<=127), min hidden count, 15 histogram bins, discard top, hand]
    const uint32_t hlo = (uint32_t)s.hist, hhi = (uint32_t)(s.hist >> 32);
    const uint32_t e0 = spread4(hlo & 0xFFFFu), e1 = spread4(hlo >> 16);
    const uint32_t e2 = spread4(hhi & 0xFFFFu), e3 = spread4(hhi >> 16);
    const uint32_t bin2 = (hlo >> 8) & 0xFFu;
    const uint32_t hand_code = (uint32_t)(s.hdr >> HDR_HAND_SH) & 0xFu;
    const uint32_t top_code = (uint32_t)(s.hdr >> HDR_TOP_SH) & 0xFu;
    const uint32_t hand_b = hand_code == HAND_NONE ? 15u : ((hand_code - 2u) & 0xFFu);
    const uint32_t top_b = (top_code - 3u) & 0xFFu;
    const int msi = (int)min_sum24 - 24;
    const uint32_t ms = (uint32_t)(msi < 127 ? msi : 127) & 0xFFu;
    o.s[0] = ms | (min_hid << 8) | (e0 << 16);
    o.s[1] = bin2 | (e1 << 8);
    o.s[2] = (e1 >> 24) | (e2 << 8);
    o.s[3] = (e2 >> 24) | (e3 << 8);
    uint32_t tail = (e3 >> 24) | (top_b << 8) | (hand_b << 16);
    if (IND) {
        const Row v = row_observed(ob);
        o.s[4] = tail | (v.w0 & 0xFF000000u);
        o.s[5] = v.w1;
        o.s[6] = v.w2;
        o.s[7] = v.w3 & 0x00FFFFFFu;
    } else {
#pragma unroll
        for (int q = 0; q < N; ++q) {
            const Row v = row_observed(s.row[q]);
            o.s[4 + 3 * q] = tail | (v.w0 & 0xFF000000u);
            o.s[5 + 3 * q] = v.w1;
            o.s[6 + 3 * q] = v.w2;
            tail = v.w3 & 0x00FFFFFFu;
        }
        o.s[4 + 3 * N] = tail;
    }
    // action mask, 26 bytes of 0/1
    if (observer_hidden) *observer_hidden = row_hidden(ob);
    if (N > SKYJO_DRAW_SHORTCUT_MAX_N || (s.hdr & HDR_PHASE)) {
        const uint32_t lb = legal_bits(row_hidden(ob), row_flags(ob), (s.hdr & HDR_PHASE) != 0);
#pragma unroll
        for (int k = 0; k < 7; ++k) o.m[k] = bits01(lb >> (4 * k));
    } else {  // draw phase: actions 24 and 25 (skyjo.py:218-221), whatever the row
#pragma unroll
        for (int k = 0; k < 6; ++k) o.m[k] = 0u;
        o.m[6] = 0x00000101u;
    }
}

// The aligned words a thread owns in its warp's slice of an output tile.  A warp's 32 rows
// are 32*D bytes (a multiple of 4), row t starts at byte t*D; every aligned word is owned by the
// row containing its first byte.  Row t owns words first .. first+count-1; word k is bytes
// [o + 4k, o + 4k + 4) of the row's stream continued by the next row's stream, so the last
// owned word needs the first bytes of the next lane's word 0 (`next0`, a warp shuffle on the
// device).  NW stream words, TAIL valid bytes in the last one.
template <int NW, int TAIL>
SKYJO_HD void stage_words(const uint32_t (&S)[NW], uint32_t next0, int row_in_tile, uint32_t (&out)[NW],
                          uint32_t &first, int &count) {
    constexpr uint32_t D = 4u * NW - 4u + TAIL;
    const uint32_t g = (uint32_t)row_in_tile * D;
    const uint32_t o = (4u - (g & 3u)) & 3u;
    first = (g + 3u) >> 2;
    const uint32_t lastw = (S[NW - 1] & (0xFFFFFFFFu >> (8 * (4 - TAIL)))) | (next0 << (8 * TAIL));
    const uint32_t beyond = next0 >> (8 * (4 - TAIL));
    const uint32_t sh = 8u * o;
#pragma unroll
    for (int k = 0; k < NW - 2; ++k) out[k] = sk_funnel_r(S[k], S[k + 1], sh);
    out[NW - 2] = sk_funnel_r(S[NW - 2], lastw, sh);
    out[NW - 1] = sk_funnel_r(lastw, beyond, sh);
    count = o < (uint32_t)TAIL ? NW : NW - 1;
}

}  // namespace skyjo
