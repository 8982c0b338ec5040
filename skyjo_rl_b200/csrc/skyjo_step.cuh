// skyjo_step.cuh -- the fused step + mask + observe kernel (one launch = one lockstep step of
// every env) and the stand-alone observe kernel.
//
// One game per thread, TILE games per CTA.  Per step a thread
//   1. loads its planes (LDG.128, coalesced over the env index),
//   2. optionally draws a uniform legal action in-kernel (random_admissible_policy.py:26-28),
//   3. applies SkyjoGame.act (reference skyjo.py:308-335): _action_draw_card (:337-374) with the
//      last-round check (:350-356), _action_place (:376-427), column removal (:431-469),
//      _evaluate_game (:477-498) and SimpleSkyjoEnv._calc_final_rewards (skyjo_env.py:293-312),
//   4. on termination installs the pre-dealt next episode (auto-reset),
//   5. stores the planes it changed and encodes the next agent's observation and action mask
//      (skyjo.py:148-224) through shared memory (skyjo_encode.cuh).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/skyjo_b200.h"
#include "skyjo_encode.cuh"
#include "skyjo_rng.cuh"
#include "skyjo_state.cuh"

namespace skyjo {

__device__ __forceinline__ uint64_t pack64(uint32_t lo, uint32_t hi) { return (uint64_t)lo | ((uint64_t)hi << 32); }

// sample one card code from a packed histogram (sampling without replacement of the
// reshuffled pile, see DESIGN.md "in-game reshuffle"); idx < total
__device__ __forceinline__ uint32_t hist_take(uint64_t &h, uint32_t idx) {
    uint32_t code = 14;
    bool found = false;
#pragma unroll
    for (uint32_t c = 0; c < 14; ++c) {
        const uint32_t cnt = hist_get(h, c);
        if (!found) {
            if (idx < cnt) {
                code = c;
                found = true;
            } else {
                idx -= cnt;
            }
        }
    }
    h -= hist_one(code);
    return code;
}

__device__ __forceinline__ uint32_t hist_total(uint64_t h) {
    uint32_t t = 0;
#pragma unroll
    for (uint32_t c = 0; c < 15; ++c) t += hist_get(h, c);
    return t;
}

// raw (unpenalised) score of a row: columns whose three cards are not all equal (skyjo.py:488-493)
__device__ __forceinline__ int row_score(uint64_t row) {
    int s = 0;
#pragma unroll
    for (int col = 0; col < 4; ++col) {
        int c0 = (int)row_code(row, 3 * col), c1 = (int)row_code(row, 3 * col + 1), c2 = (int)row_code(row, 3 * col + 2);
        if (!(c0 == c1 && c1 == c2)) s += c0 + c1 + c2 - 6;
    }
    return s;
}

// numpy's add.reduce order for a float64 vector of N <= 12 entries
template <int N>
__device__ __forceinline__ double np_sum(const double (&a)[N]) {
    if constexpr (N < 8) {
        double r = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) r = __dadd_rn(r, a[i]);
        return r;
    } else {
        double r = __dadd_rn(__dadd_rn(__dadd_rn(a[0], a[1]), __dadd_rn(a[2], a[3])),
                             __dadd_rn(__dadd_rn(a[4], a[5]), __dadd_rn(a[6], a[7])));
#pragma unroll
        for (int i = 8; i < N; ++i) r = __dadd_rn(r, a[i]);
        return r;
    }
}

__device__ __forceinline__ int load_action(const void *actions, int dtype, long long e) {
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

template <int N, bool IND, bool POLICY>
__global__ void __launch_bounds__(TILE) step_kernel(const StepParams p) {
    constexpr int NP = (N + 1) / 2;
    constexpr int D = IND ? 31 : 19 + 12 * N;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *s_obs = smem;
    uint8_t *s_mask = smem + TILE * D;  // TILE*D is a multiple of 16
    __shared__ unsigned long long s_stats[NUM_STATS];

    const int tid = threadIdx.x;
    const long long tile0 = (long long)blockIdx.x * TILE;
    const long long e = tile0 + tid;
    const bool valid = e < p.B;
    if (tid < NUM_STATS) s_stats[tid] = 0ull;

    // ---- 1. load state ----------------------------------------------------------------
    const uint4 P0 = p.st.planes[e];
    uint64_t rows[N];
#pragma unroll
    for (int k = 0; k < NP; ++k) {
        const uint4 T = p.st.planes[(long long)(1 + k) * p.Bpad + e];
        rows[2 * k] = pack64(T.x, T.y);
        if (2 * k + 1 < N) rows[2 * k + 1] = pack64(T.z, T.w);
    }
    int action = 0;
    if (!POLICY && valid) action = load_action(p.actions, p.action_dtype, e);
    __syncthreads();  // s_stats zeroed

    uint64_t hdr = pack64(P0.x, P0.y);
    uint64_t hist = pack64(P0.z, P0.w);
    const int cur = (int)(hdr >> HDR_CUR_SH) & 0xF;
    const bool place_phase = (hdr & HDR_PHASE) != 0;
    uint64_t row = rows[0];
#pragma unroll
    for (int q = 1; q < N; ++q)
        if (q == cur) row = rows[q];
    uint32_t hidden = (uint32_t)(row >> 48) & 0xFFFu;
    uint32_t flags = (uint32_t)(row >> 60);

    uint32_t dirty_planes = 0;  // bit k: table plane k must be written back
    int done_code = SKYJO_RUNNING;
    int act_class = -1;  // 0: 24, 1: 25, 2: swap, 3: flip
    const bool frozen = (hdr & HDR_TERMINATED) != 0;

    if (valid && (hdr & HDR_DIRTY)) {
        // the previous step ended an episode: its rewards have been consumed
#pragma unroll
        for (int q = 0; q < N; ++q) p.reward[e * N + q] = 0.0;
        hdr &= ~HDR_DIRTY;
    }

    if (valid && frozen) {
        done_code = SKYJO_DONE_GAME_OVER;  // skyjo.py:316-321: playing a finished game returns True
    } else if (valid) {
        const uint32_t legal = legal_bits(hidden, flags, place_phase);
        if (POLICY) action = policy_pick(p.seed, p.first_env + (unsigned long long)e, p.t, legal);
        const bool is_legal = action >= 0 && action < 26 && ((legal >> action) & 1u);
        uint32_t step = (uint32_t)(hdr & HDR_STEP_MASK);
        if (!is_legal) {
            // TerminateIllegalWrapper(illegal_reward=-1), skyjo_env.py:23
            done_code = SKYJO_DONE_ILLEGAL;
#pragma unroll
            for (int q = 0; q < N; ++q) p.reward[e * N + q] = (q == cur) ? -1.0 : 0.0;
            atomicAdd(&s_stats[SKYJO_STAT_ILLEGAL], 1ull);
        } else if (!place_phase) {
            act_class = action - 24;
            if (hidden == 0) {
                // ---- game over (skyjo.py:350-356): score, penalty, rewards ---------------
                done_code = SKYJO_DONE_GAME_OVER;
                int raw[N];
                int mn = 1 << 30, refunds = 0;
#pragma unroll
                for (int q = 0; q < N; ++q) {
                    const uint64_t r = (q == cur) ? row : rows[q];
                    raw[q] = row_score(r);
                    mn = min(mn, raw[q]);
                    refunds += __popc((uint32_t)(r >> 60));
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
                    if (penalised && q == cur) score[q] = __dmul_rn(score[q], p.score_penalty);
                }
                // skyjo_env.py:307-311
                const double mean = __ddiv_rn(np_sum<N>(score), (double)N);
                int winner = 0;
                double best = score[0];
#pragma unroll
                for (int q = 0; q < N; ++q) {
                    double r = __dadd_rn(__dadd_rn(-score[q], mean), p.mean_reward);
                    if (p.reward_refunded != 0.0) {
                        const uint64_t rq = (q == cur) ? row : rows[q];
                        r = __dadd_rn(r, __dmul_rn((double)__popc((uint32_t)(rq >> 60)), p.reward_refunded));
                    }
                    p.reward[e * N + q] = r;
                    p.final_score[e * N + q] = score[q];
                    if (score[q] < best) {
                        best = score[q];
                        winner = q;
                    }
                }
                int raw_sum = 0;
#pragma unroll
                for (int q = 0; q < N; ++q) raw_sum += raw[q];
                atomicAdd(&s_stats[SKYJO_STAT_EPISODES], 1ull);
                atomicAdd(&s_stats[SKYJO_STAT_EPISODE_STEPS], (unsigned long long)(step + 1));
                atomicAdd(&s_stats[SKYJO_STAT_SCORE_RAW_SUM], (unsigned long long)(long long)raw_sum);
                atomicAdd(&s_stats[SKYJO_STAT_WINNER_RAW_SUM], (unsigned long long)(long long)mn);
                atomicAdd(&s_stats[SKYJO_STAT_FINISHER_RAW_SUM], (unsigned long long)(long long)fin_raw);
                if (penalised) {
                    atomicAdd(&s_stats[SKYJO_STAT_PENALISED], 1ull);
                    atomicAdd(&s_stats[SKYJO_STAT_PENALISED_RAW_SUM], (unsigned long long)(long long)fin_raw);
                }
                atomicAdd(&s_stats[SKYJO_STAT_REFUNDS], (unsigned long long)refunds);
                if (((hdr >> HDR_STARTER_SH) & 0xF) == 0) atomicAdd(&s_stats[SKYJO_STAT_STARTER_SEAT0], 1ull);
                atomicAdd(&s_stats[SKYJO_STAT_WINS_SEAT0 + winner], 1ull);
            } else {
                uint32_t code;
                if (action == 24) {
                    // ---- draw from the draw pile (skyjo.py:359-366) ----------------------
                    uint32_t n_draw = (uint32_t)(hdr >> HDR_NDRAW_SH) & 0xFFu;
                    const uint32_t slot = (hdr & HDR_SLOT) ? 1u : 0u;
                    uint8_t *prow = p.st.pile + ((long long)slot * p.Bpad + e) * PILE_ROW;
                    if (n_draw == 0) {
                        // reshuffle the whole discard pile into a new draw pile (:361-365)
                        uint64_t dh = hist;
                        if (!IND) {  // direct-mode hist also counts the open table cards
#pragma unroll
                            for (int q = 0; q < N; ++q) {
                                const uint64_t r = (q == cur) ? row : rows[q];
                                const uint32_t open = ~((uint32_t)(r >> 48) | cols_to_slots((uint32_t)(r >> 60))) & 0xFFFu;
                                for (uint32_t s = 0; s < 12; ++s)
                                    if ((open >> s) & 1u) dh -= hist_one(row_code(r, s));
                            }
                        }
                        const uint32_t total = hist_total(dh);
                        uint32_t ep = p.st.episode[e] - 1u;
                        if ((ep & 15u) != ((uint32_t)(hdr >> HDR_EPLO_SH) & 15u)) ep -= 1u;
                        const uint32_t q8 = (uint32_t)(hdr >> HDR_Q_SH) & 0xFFu;
                        const unsigned long long genv = p.first_env + (unsigned long long)e;
                        U4 r0 = rng_block(p.seed, genv, PURPOSE_RESHUFFLE, ep, (q8 << 16) | total);
                        uint64_t left = dh;
                        const uint32_t e0 = hist_take(left, bounded(r0.x, total));
                        hist = hist - dh + hist_one(e0);  // new discard pile = [e0]
                        hdr = (hdr & ~((0xFFull << HDR_TOP_SH) | (0xFFull << HDR_Q_SH))) |
                              ((uint64_t)(e0 + 1u) << HDR_TOP_SH) | ((uint64_t)((q8 + 1u) & 0xFFu) << HDR_Q_SH) | HDR_LAZY;
                        n_draw = total - 1u;
                        U4 r1 = rng_block(p.seed, genv, PURPOSE_RESHUFFLE, ep, (q8 << 16) | n_draw);
                        code = hist_take(left, bounded(r1.x, n_draw));
                        *reinterpret_cast<uint64_t *>(prow) = left;
                        atomicAdd(&s_stats[SKYJO_STAT_RESHUFFLES], 1ull);
                    } else if (hdr & HDR_LAZY) {
                        uint64_t left = *reinterpret_cast<const uint64_t *>(prow);
                        uint32_t ep = p.st.episode[e] - 1u;
                        if ((ep & 15u) != ((uint32_t)(hdr >> HDR_EPLO_SH) & 15u)) ep -= 1u;
                        const uint32_t q8 = ((uint32_t)(hdr >> HDR_Q_SH) - 1u) & 0xFFu;
                        U4 r1 = rng_block(p.seed, p.first_env + (unsigned long long)e, PURPOSE_RESHUFFLE, ep,
                                          (q8 << 16) | n_draw);
                        code = hist_take(left, bounded(r1.x, n_draw));
                        *reinterpret_cast<uint64_t *>(prow) = left;
                    } else {
                        code = prow[n_draw - 1u];
                    }
                    n_draw -= 1u;
                    hdr = (hdr & ~(0xFFull << HDR_NDRAW_SH)) | ((uint64_t)n_draw << HDR_NDRAW_SH);
                } else {
                    // ---- take the discard top (skyjo.py:370) ---------------------------------
                    const uint32_t top = (uint32_t)(hdr >> HDR_TOP_SH) & 0xFu;
                    const uint32_t second = (uint32_t)(hdr >> HDR_SECOND_SH) & 0xFu;
                    code = top - 1u;
                    hist -= hist_one(code);
                    hdr = (hdr & ~(0xFFull << HDR_TOP_SH)) | ((uint64_t)second << HDR_TOP_SH);
                }
                hdr = (hdr & ~(0xFull << HDR_HAND_SH)) | ((uint64_t)code << HDR_HAND_SH) | HDR_PHASE;
            }
        } else {
            // ---- place (skyjo.py:376-427) -------------------------------------------------
            const uint32_t hand = (uint32_t)(hdr >> HDR_HAND_SH) & 0xFu;
            uint32_t top = (uint32_t)(hdr >> HDR_TOP_SH) & 0xFu;
            uint32_t second;
            uint32_t s;
            if (action < 12) {  // :389-395 swap
                act_class = 2;
                s = (uint32_t)action;
                const uint32_t old = row_code(row, s);
                const bool was_hidden = (hidden >> s) & 1u;
                second = top;
                top = old + 1u;
                if (IND || was_hidden) hist += hist_one(old);
                if (!IND) hist += hist_one(hand);
                row = row_set_code(row, s, hand);
            } else {  // :396-404 discard the hand card and reveal
                act_class = 3;
                s = (uint32_t)action - 12u;
                second = top;
                top = hand + 1u;
                hist += hist_one(hand);
                if (!IND) hist += hist_one(row_code(row, s));
            }
            hidden &= ~(1u << s);
            // column removal (:431-469): only the touched column can newly qualify
            const uint32_t col = s / 3u;
            const uint32_t c0 = row_code(row, 3 * col), c1 = row_code(row, 3 * col + 1), c2 = row_code(row, 3 * col + 2);
            if (c0 == c1 && c1 == c2 && ((hidden >> (3 * col)) & 7u) == 0 && !((flags >> col) & 1u)) {
                if (!IND) hist -= 3ull * hist_one(c0);
                hist += 3ull * hist_one(2u);  // three zeros go to the discard pile (:454-458)
                top = 3u;
                second = 3u;
                flags |= 1u << col;
                row = (row & ~(0xFFFull << (12 * col))) | (0x222ull << (12 * col));
            }
            row = (row & 0x0000FFFFFFFFFFFFull) | ((uint64_t)hidden << 48) | ((uint64_t)flags << 60);
            const int nxt = (cur + 1 == N) ? 0 : cur + 1;
            hdr = (hdr & ~((0xFull << HDR_CUR_SH) | HDR_PHASE | (0xFFFull << HDR_HAND_SH))) |
                  ((uint64_t)nxt << HDR_CUR_SH) | ((uint64_t)HAND_NONE << HDR_HAND_SH) |
                  ((uint64_t)top << HDR_TOP_SH) | ((uint64_t)second << HDR_SECOND_SH);
#pragma unroll
            for (int q = 0; q < N; ++q)
                if (q == cur) rows[q] = row;
            dirty_planes |= 1u << (cur >> 1);
        }
        if (done_code == SKYJO_RUNNING) {
            step = min(step + 1u, 0xFFFFu);
            hdr = (hdr & ~HDR_STEP_MASK) | step;
            if (p.max_steps > 0 && step >= (uint32_t)p.max_steps) {
                done_code = SKYJO_DONE_TRUNCATED;
#pragma unroll
                for (int q = 0; q < N; ++q) p.reward[e * N + q] = 0.0;
                atomicAdd(&s_stats[SKYJO_STAT_TRUNCATED], 1ull);
            }
        }
        // ---- 4. episode end: install the pre-dealt next episode, or freeze ---------------
        if (done_code != SKYJO_RUNNING) {
            bool installed = false;
            if (p.auto_reset) {
                const uint4 Q0 = p.st.next_planes[e];
                const uint32_t want = ((uint32_t)(hdr >> HDR_EPLO_SH) + 1u) & 15u;
                if ((Q0.y & 15u) == want) {
                    const uint32_t old_slot = (hdr & HDR_SLOT) ? 1u : 0u;
                    hdr = pack64(Q0.x, Q0.y) | HDR_DIRTY;
                    hist = pack64(Q0.z, Q0.w);
#pragma unroll
                    for (int k = 0; k < NP; ++k) {
                        const uint4 T = p.st.next_planes[(long long)(1 + k) * p.Bpad + e];
                        rows[2 * k] = pack64(T.x, T.y);
                        if (2 * k + 1 < N) rows[2 * k + 1] = pack64(T.z, T.w);
                    }
                    dirty_planes = (1u << NP) - 1u;
                    p.st.needs_deal[e] = (uint8_t)(1u | (old_slot << 1));
                    installed = true;
                } else {
                    atomicOr(p.st.errflag, ERR_NEXT_NOT_READY);
                }
            }
            if (!installed) hdr |= HDR_TERMINATED;
        }
    }

    // ---- action-class statistics (warp ballots, one shared atomic per class and warp) ------
    {
        const unsigned b0 = __ballot_sync(0xFFFFFFFFu, act_class == 0);
        const unsigned b1 = __ballot_sync(0xFFFFFFFFu, act_class == 1);
        const unsigned b2 = __ballot_sync(0xFFFFFFFFu, act_class == 2);
        const unsigned b3 = __ballot_sync(0xFFFFFFFFu, act_class == 3);
        if ((tid & 31) == 0) {
            const unsigned n0 = __popc(b0), n1 = __popc(b1), n2 = __popc(b2), n3 = __popc(b3);
            if (n0) atomicAdd(&s_stats[SKYJO_STAT_ACT_DRAW_PILE], (unsigned long long)n0);
            if (n1) atomicAdd(&s_stats[SKYJO_STAT_ACT_TAKE_DISCARD], (unsigned long long)n1);
            if (n2) atomicAdd(&s_stats[SKYJO_STAT_ACT_SWAP], (unsigned long long)n2);
            if (n3) atomicAdd(&s_stats[SKYJO_STAT_ACT_FLIP], (unsigned long long)n3);
            if (n0 + n1 + n2 + n3) atomicAdd(&s_stats[SKYJO_STAT_STEPS], (unsigned long long)(n0 + n1 + n2 + n3));
        }
    }

    // ---- 5. write back state ------------------------------------------------------------
    if (valid) {
        p.st.planes[e] = make_uint4((uint32_t)hdr, (uint32_t)(hdr >> 32), (uint32_t)hist, (uint32_t)(hist >> 32));
#pragma unroll
        for (int k = 0; k < NP; ++k) {
            if ((dirty_planes >> k) & 1u) {
                uint4 T;
                T.x = (uint32_t)rows[2 * k];
                T.y = (uint32_t)(rows[2 * k] >> 32);
                if (2 * k + 1 < N) {
                    T.z = (uint32_t)rows[2 * k + 1];
                    T.w = (uint32_t)(rows[2 * k + 1] >> 32);
                } else {
                    T.z = 0;
                    T.w = 0;
                }
                p.st.planes[(long long)(1 + k) * p.Bpad + e] = T;
            }
        }
        p.agent[e] = (int8_t)((hdr >> HDR_CUR_SH) & 0xF);
        p.done[e] = (uint8_t)done_code;
    }

    // ---- 6. observation + mask of the next agent -----------------------------------------
    encode_rows<N, IND>(rows, hdr, hist, (int)(hdr >> HDR_CUR_SH) & 0xF, s_obs, s_mask, tid);
    fence_async_smem();
    __syncthreads();
    const long long left = p.B - tile0;
    const uint32_t n_env = left >= TILE ? TILE : (uint32_t)left;
    const bool bulk = p.bulk_ok && n_env == TILE;
    store_tile(s_obs, p.obs + tile0 * D, n_env * D, bulk, tid);
    store_tile(s_mask, p.mask + tile0 * 26, n_env * 26u, bulk, tid);
    if (tid < NUM_STATS) {
        const unsigned long long v = s_stats[tid];
        if (v) atomicAdd(&p.st.stats[(blockIdx.x % STAT_SLOTS) * NUM_STATS + tid], v);
    }
    if (bulk) bulk_commit_and_wait(tid);
}

// Stand-alone observe (SimpleSkyjoEnv.observe, skyjo_env.py:199-214): encodes the view of
// `agent` (or of each env's agent_selection when agent < 0) without touching the state.
// With reset_outputs it also publishes agent / done = 0 / reward = 0 (after a reset).
template <int N, bool IND>
__global__ void __launch_bounds__(TILE) observe_kernel(const StepParams p, int agent, int8_t *obs_out, int8_t *mask_out,
                                                       int reset_outputs, int bulk_ok) {
    constexpr int NP = (N + 1) / 2;
    constexpr int D = IND ? 31 : 19 + 12 * N;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *s_obs = smem;
    uint8_t *s_mask = smem + TILE * D;
    const int tid = threadIdx.x;
    const long long tile0 = (long long)blockIdx.x * TILE;
    const long long e = tile0 + tid;
    const uint4 P0 = p.st.planes[e];
    uint64_t rows[N];
#pragma unroll
    for (int k = 0; k < NP; ++k) {
        const uint4 T = p.st.planes[(long long)(1 + k) * p.Bpad + e];
        rows[2 * k] = pack64(T.x, T.y);
        if (2 * k + 1 < N) rows[2 * k + 1] = pack64(T.z, T.w);
    }
    const uint64_t hdr = pack64(P0.x, P0.y), hist = pack64(P0.z, P0.w);
    const int cur = (int)(hdr >> HDR_CUR_SH) & 0xF;
    if (reset_outputs && e < p.B) {
        p.agent[e] = (int8_t)cur;
        p.done[e] = 0;
#pragma unroll
        for (int q = 0; q < N; ++q) {
            p.reward[e * N + q] = 0.0;
            p.final_score[e * N + q] = 0.0;
        }
    }
    encode_rows<N, IND>(rows, hdr, hist, agent < 0 ? cur : agent, s_obs, s_mask, tid);
    fence_async_smem();
    __syncthreads();
    const long long left = p.B - tile0;
    const uint32_t n_env = left >= TILE ? TILE : (uint32_t)left;
    const bool bulk = bulk_ok && n_env == TILE;
    store_tile(s_obs, obs_out + tile0 * D, n_env * D, bulk, tid);
    store_tile(s_mask, mask_out + tile0 * 26, n_env * 26u, bulk, tid);
    if (bulk) bulk_commit_and_wait(tid);
}

// launchers implemented per player count in skyjo_step_inst.cu
typedef cudaError_t (*step_launch_fn)(const StepParams &, bool indirect, bool policy, cudaStream_t);
typedef cudaError_t (*observe_launch_fn)(const StepParams &, bool indirect, int agent, int8_t *, int8_t *, int, int,
                                         cudaStream_t);

}  // namespace skyjo
