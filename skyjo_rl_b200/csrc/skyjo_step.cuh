// skyjo_step.cuh -- the fused step + mask + observe kernel (one launch = one lockstep step of
// every env) and the stand-alone observe kernel.
//
// One game per thread, TILE games per CTA.  Per step a thread
//   1. loads its 1 + N planes (LDG.128, coalesced over the env index),
//   2. runs env_step (skyjo_core.cuh): optional in-kernel uniform legal policy, SkyjoGame.act,
//      end-of-game scoring and rewards, auto-reset install of the pre-dealt next episode,
//   3. stores plane 0 and the one row it changed (STG.128),
//   4. encodes the next agent's observation and action mask as word streams and stages them
//      through shared memory into one TMA bulk store per tensor (skyjo_encode.cuh).
// Statistics are reduced per warp (REDUX) and per CTA (32-bit shared atomics) before one
// 64-bit global atomic per non-zero entry.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/skyjo_b200.h"
#include "skyjo_core.cuh"
#include "skyjo_encode.cuh"
#include "skyjo_rng.cuh"
#include "skyjo_state.cuh"

namespace skyjo {

// resident warps per SM the kernels are compiled for (caps registers: 2048 / warps per thread).
// Measured on B200 (tools/variants.py, next-step reset, 2^20 / 2^22 envs): N <= 4: 32 warps (64
// registers) beat 28 / 24; N = 5: 28 (-2.7 % vs 32, which spills); N = 6: 28; N = 7: 24; N = 8: 22 (88 registers; 28: +7 %, 24: +0.6 %,
// 20: +4.5 %); N >= 9: 12 (16 / 20 / 24: +4 / +3 / +23 %).  The rollout kernel keeps the state live
// across its steps and likes one notch fewer warps at N = 8.
#ifndef SKYJO_STEP_WARPS_SMALL
#define SKYJO_STEP_WARPS_SMALL 32
#endif
#ifndef SKYJO_STEP_WARPS_MID
#define SKYJO_STEP_WARPS_MID(N) ((N) == 6 ? 28 : ((N) == 7 ? 24 : 22))
#endif
#ifndef SKYJO_STEP_WARPS_LARGE
#define SKYJO_STEP_WARPS_LARGE 12
#endif
#ifndef SKYJO_ROLLOUT_WARPS_SMALL
#define SKYJO_ROLLOUT_WARPS_SMALL 32
#endif
#ifndef SKYJO_ROLLOUT_WARPS_MID
#define SKYJO_ROLLOUT_WARPS_MID(N) ((N) == 6 ? 28 : ((N) == 7 ? 24 : 20))
#endif
// player count from which the warp assist (below) pays: at N <= 5 scoring N rows is short and
// reshuffles do not occur, so its per-step test costs more than it saves (measured: N=4 +2 %,
// N=5 +-0, N=8 -17 %)
#ifndef SKYJO_ASSIST_MIN_N
#define SKYJO_ASSIST_MIN_N 6
#endif
#ifndef SKYJO_STEP_WARPS_5
#define SKYJO_STEP_WARPS_5 28
#endif
#define STEP_WARPS_PER_SM(N) \
    ((N) <= 4 ? SKYJO_STEP_WARPS_SMALL  \
              : ((N) == 5 ? SKYJO_STEP_WARPS_5 : ((N) <= 8 ? SKYJO_STEP_WARPS_MID(N) : SKYJO_STEP_WARPS_LARGE)))
#define ROLLOUT_WARPS_PER_SM(N) \
    ((N) <= 5 ? SKYJO_ROLLOUT_WARPS_SMALL : ((N) <= 8 ? SKYJO_ROLLOUT_WARPS_MID(N) : SKYJO_STEP_WARPS_LARGE))
#define STEP_MIN_CTAS(N) ((STEP_WARPS_PER_SM(N) * 32) / TILE)
#define ROLLOUT_MIN_CTAS(N) ((ROLLOUT_WARPS_PER_SM(N) * 32) / TILE)

constexpr int WARPS = TILE / 32;

// ---- L2 residency hints -------------------------------------------------------------------------
// The live planes of 2^20 4-player envs are 84 MB, the outputs of one step 100 MB, the L2 126 MB.
// Without hints the output stream evicts the planes between two steps and every step re-reads
// them from DRAM.  SKYJO_L2_KEEP marks plane loads / stores evict_last, SKYJO_L2_STREAM marks the
// output stores evict_first, so that the state stays L2-resident across steps when it fits.
#ifndef SKYJO_L2_KEEP
#define SKYJO_L2_KEEP 0
#endif
#ifndef SKYJO_L2_STREAM
#define SKYJO_L2_STREAM 0
#endif
__device__ __forceinline__ uint64_t l2_policy_keep() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t l2_policy_stream() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ U128 ld128_hint(const U128 *p, uint64_t pol) {
    U128 v;
    asm volatile("ld.global.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void st128_hint(U128 *p, uint32_t x, uint32_t y, uint32_t z, uint32_t w, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "r"(x), "r"(y), "r"(z),
                 "r"(w), "l"(pol)
                 : "memory");
}
template <int N>
__device__ __forceinline__ void load_env_dev(const U128 *planes, long long Bpad, long long e, Env<N> &s) {
#if SKYJO_L2_KEEP
    const uint64_t pol = l2_policy_keep();
    const U128 P0 = ld128_hint(planes + e, pol);
    s.hdr = pack64(P0.x, P0.y);
    s.hist = pack64(P0.z, P0.w);
#pragma unroll
    for (int q = 0; q < N; ++q) {
        const U128 T = ld128_hint(planes + (long long)(1 + q) * Bpad + e, pol);
        s.row[q].w0 = T.x;
        s.row[q].w1 = T.y;
        s.row[q].w2 = T.z;
        s.row[q].w3 = T.w;
    }
#else
    load_env<N>(planes, Bpad, e, s);
#endif
}
template <int N>
__device__ __forceinline__ void store_env_dev(U128 *planes, long long Bpad, long long e, const Env<N> &s,
                                              uint32_t dirty_rows, uint32_t pf_new) {
#if SKYJO_L2_KEEP
    const uint64_t pol = l2_policy_keep();
    const uint64_t hdr = pf_new == 0xFFFFFFFFu ? s.hdr : hdr_pf_set(s.hdr, pf_new);
    st128_hint(planes + e, (uint32_t)hdr, (uint32_t)(hdr >> 32), (uint32_t)s.hist, (uint32_t)(s.hist >> 32), pol);
#pragma unroll
    for (int q = 0; q < N; ++q)
        if ((dirty_rows >> q) & 1u)
            st128_hint(planes + (long long)(1 + q) * Bpad + e, s.row[q].w0, s.row[q].w1, s.row[q].w2, s.row[q].w3, pol);
#else
    store_env<N>(planes, Bpad, e, s, dirty_rows, pf_new);
#endif
}

// True when the NEXT step of this env will install its pre-dealt episode (so its planes are worth
// pulling into L2 now).  Same-step reset: the agent on turn is about to draw with no hidden card
// left, which ends the game (skyjo.py:350-356).  Next-step reset: the episode has just ended.
__device__ __forceinline__ bool install_next_step(int auto_reset, uint64_t hdr, uint32_t next_hidden) {
    if (auto_reset == 2) return (hdr & HDR_TERMINATED) != 0;
    return auto_reset && next_hidden == 0u && !(hdr & (HDR_PHASE | HDR_TERMINATED));
}

// 1: below the assist threshold, games that end in the phase-locked mode are scored by the warp after env_step
// (warp_score_deferred).  Bit-exact (the GPU parity suites pass with it) but measured slower, see there: off.
#ifndef SKYJO_DEFER_SCORING
#define SKYJO_DEFER_SCORING 0
#endif
// ---- warp assist for the rare, long events ---------------------------------------------------------
// End-of-game scoring (skyjo.py:477-498 + skyjo_env.py:293-312: N rows of 12 cards, float64 rewards)
// and the discard-only histogram of an in-game reshuffle in direct mode (12 N table slots) hit about
// one env in 130 / 330 per step, i.e. one or two lanes of a third of the warps, and their scalar code
// is 350 - 1500 instructions that the other thirty lanes sit through (14 % of the N=4 kernel's
// instructions, 42 % at N=8: profiles/).  Here the warp does them together: the env's rows go
// through `scratch` (>= 16 N bytes of the warp's staging tile, not yet in use) to lanes 0..N-1, one
// row per lane, and reductions bring the results back to the owning lane, which hands them to
// env_step<.., ASSIST = true>.  The arithmetic (order of the float64 sums included) is that of the
// scalar code in skyjo_core.cuh, which the host build and tests/hostsim keep using.
// End-of-game scoring by the warp (skyjo.py:477-498 + skyjo_env.py:293-312), called by all 32 lanes: lane q < N
// (`act`) holds row q of the env whose global index is eL, `fin` is its finisher; rewards and final scores are
// stored by the N lanes, the reduced statistics land in `as` of the lane with `owner`.  Same arithmetic, float64
// summation order included, as score_game (skyjo_core.cuh).
template <int N>
__device__ __forceinline__ void warp_score_row(const StepParams &p, const Row &mine, bool act, int fin, long long eL,
                                               int lane, bool owner, Assist &as) {
    constexpr unsigned FULL = 0xFFFFFFFFu;
    uint32_t v[3];
    row_cards(mine, v);  // hidden cards count at their true value (skyjo.py:488-493)
    const int raw = score12(v);
    const int ref = (int)sk_popc(row_flags(mine));
    const int mn = __reduce_min_sync(FULL, act ? raw : 0x7FFFFFFF);
    const int raw_sum = __reduce_add_sync(FULL, act ? raw : 0);
    const int refunds = __reduce_add_sync(FULL, act ? ref : 0);
    const int fin_raw = __shfl_sync(FULL, raw, fin);
    const bool penalised = mn != fin_raw;  // skyjo.py:496
    double sc = (double)raw;
    if (penalised && lane == fin) sc = sk_dmul(sc, p.score_penalty);
    double a[N];
#pragma unroll
    for (int q = 0; q < N; ++q) a[q] = __shfl_sync(FULL, sc, q);
    const double mean = sk_ddiv(np_sum<N>(a), (double)N);  // skyjo_env.py:307-311
    int winner = 0;
    double best = a[0];
#pragma unroll
    for (int q = 1; q < N; ++q)
        if (a[q] < best) {
            best = a[q];
            winner = q;
        }
    if (act) {
        double r = sk_dadd(sk_dadd(-sc, mean), p.mean_reward);
        if (p.reward_refunded != 0.0) r = sk_dadd(r, sk_dmul((double)ref, p.reward_refunded));
        p.reward[eL * N + lane] = r;
        p.final_score[eL * N + lane] = sc;
    }
    if (owner) {
        as.scored = 1;
        as.raw_sum = raw_sum;
        as.winner_raw = mn;
        as.fin_raw = fin_raw;
        as.penalised = penalised ? 1 : 0;
        as.refunds = refunds;
        as.winner = winner;
    }
}

// Deferred scoring below the assist threshold (phase-locked reset mode only): env_step<.., DEFER> left
// oc.scored == 2 on the lanes whose game just ended and their rows in `s`.  One such lane after the other hands
// its N rows through `scratch` (the warp's staging tile, not yet in use) to lanes 0..N-1, which score one row each
// -- about 120 warp-instructions per ended game instead of 350 executed by a single active lane (39 % of the
// warps of a draw slot hold such a lane at N = 4).  No test runs before env_step and nothing stays live across it.
// Measured on B200 (2^20 envs, next-step reset): step kernel 44.13 against 43.91 us at N = 4, 38.84 against 38.30
// at N = 2; rollout kernel 31.46 against 30.15 us per step (its spills grow from 24 to 80 bytes).  The issue slots
// of the one-lane block are slots the SM had free (issue 66 % busy): removing them buys nothing, the ballot and the
// second warp-wide phase cost a little.  Kept behind SKYJO_DEFER_SCORING, off.
template <int N>
__device__ __forceinline__ void warp_score_deferred(const StepParams &p, const Env<N> &s, long long e, int lane,
                                                    uint8_t *scratch, Outcome &oc) {
    constexpr unsigned FULL = 0xFFFFFFFFu;
    unsigned need = __ballot_sync(FULL, oc.scored == 2);
    uint4 *sc4 = reinterpret_cast<uint4 *>(scratch);
    const int cur = (int)(s.hdr >> HDR_CUR_SH) & 0xF;  // the finisher: a game-ending draw does not advance the turn
    while (need) {  // warp-uniform
        const int L = __ffs(need) - 1;
        need &= need - 1u;
        if (lane == L) {
#pragma unroll
            for (int k = 0; k < N; ++k) sc4[k] = make_uint4(s.row[k].w0, s.row[k].w1, s.row[k].w2, s.row[k].w3);
        }
        __syncwarp();
        const bool act = lane < N;
        const uint4 t = sc4[act ? lane : 0];
        __syncwarp();
        const Row mine = {t.x, t.y, t.z, t.w};
        Assist as;
        warp_score_row<N>(p, mine, act, __shfl_sync(FULL, cur, L), e - lane + L, lane, lane == L, as);
        if (lane == L) {
            oc.scored = 1;
            oc.raw_sum = as.raw_sum;
            oc.winner_raw = as.winner_raw;
            oc.fin_raw = as.fin_raw;
            oc.penalised = as.penalised;
            oc.refunds = as.refunds;
            oc.winner = as.winner;
        }
    }
}

template <int N, bool IND, bool POLICY>
__device__ __forceinline__ void warp_assist(const StepParams &p, const Env<N> &s, long long e, bool valid, int action,
                                            uint32_t policy_rnd, int lane, uint8_t *scratch, Assist &as) {
    constexpr unsigned FULL = 0xFFFFFFFFu;
    as.scored = 0;
    as.has_dh = 0;
    const uint64_t hdr = s.hdr;
    const int cur = (int)(hdr >> HDR_CUR_SH) & 0xF;
    uint32_t w0c = s.row[0].w0;
#pragma unroll
    for (int q = 1; q < N; ++q)
        if (q == cur) w0c = s.row[q].w0;
    const uint32_t hidden = w0c & 0xFFFu;
    // the conditions under which env_step reaches the two events (same-phase test, legality, skyjo.py:350, :361)
    const bool draws = valid && !(hdr & (HDR_PHASE | HDR_TERMINATED)) && (POLICY || action == 24 || action == 25);
    const bool over = draws && hidden == 0u;
    const bool from_pile = POLICY ? bounded(policy_rnd, 2u) == 0u : action == 24;
    const bool resh = !IND && draws && hidden != 0u && from_pile && ((uint32_t)(hdr >> HDR_NDRAW_SH) & 0xFFu) == 0u;
    unsigned need = __ballot_sync(FULL, over || resh);
    uint4 *sc4 = reinterpret_cast<uint4 *>(scratch);
    while (need) {  // warp-uniform
        const int L = __ffs(need) - 1;
        need &= need - 1u;
        if (lane == L) {
#pragma unroll
            for (int k = 0; k < N; ++k) sc4[k] = make_uint4(s.row[k].w0, s.row[k].w1, s.row[k].w2, s.row[k].w3);
        }
        __syncwarp();
        const bool act = lane < N;
        const uint4 t = sc4[act ? lane : 0];
        __syncwarp();
        const Row mine = {t.x, t.y, t.z, t.w};
        const int is_over = __shfl_sync(FULL, (int)over, L);
        if (is_over) {
            const int curL = __shfl_sync(FULL, cur, L);
            warp_score_row<N>(p, mine, act, curL, e - lane + L, lane, lane == L, as);
        } else {
            // open table cards of my row, as histogram increments (skyjo.py:241-246 count_players_cards)
            uint64_t c = 0;
            if (act) {
                const uint32_t open = ~(row_hidden(mine) | cols_to_slots(row_flags(mine))) & 0xFFFu;
#pragma unroll
                for (uint32_t sl = 0; sl < 12; ++sl)
                    if ((open >> sl) & 1u) c += hist_one((row_byte(mine, sl) + 2u) & 0xFFu);
            }
#pragma unroll
            for (int off = 8; off >= 1; off >>= 1) c += __shfl_down_sync(FULL, c, off);  // N <= 12 lanes hold data
            const uint64_t tot = __shfl_sync(FULL, c, 0);
            if (lane == L) {
                as.has_dh = 1;
                as.dh = s.hist - tot;
            }
        }
    }
}

// Stores one warp's 32-row slice of an output tile: one TMA bulk store issued by lane 0 when the
// slice is complete and 16-byte aligned, a byte loop otherwise (ragged last tile).
__device__ __forceinline__ void store_warp_slice(const uint8_t *s_src, int8_t *g_dst, uint32_t row_bytes,
                                                 uint32_t rows_valid, bool bulk_ok, int lane) {
    if (bulk_ok && rows_valid == 32u) {
        if (lane == 0) {
            const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(s_src);
#if SKYJO_L2_STREAM
            const uint64_t pol = l2_policy_stream();
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;\n" ::"l"(g_dst),
                         "r"(saddr), "r"(32u * row_bytes), "l"(pol)
                         : "memory");
#else
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(g_dst),
                         "r"(saddr), "r"(32u * row_bytes)
                         : "memory");
#endif
        }
    } else {
        for (uint32_t i = lane; i < rows_valid * row_bytes; i += 32u) g_dst[i] = (int8_t)s_src[i];
    }
}

template <int N, bool IND, bool POLICY>
__global__ void __launch_bounds__(TILE, STEP_MIN_CTAS(N)) step_kernel(const __grid_constant__ StepParams p) {
    using OW = ObsWords<N, IND>;
    constexpr int D = OW::D;
    extern __shared__ __align__(128) uint8_t smem[];
    uint32_t *s_obs = reinterpret_cast<uint32_t *>(smem);
    uint32_t *s_mask = reinterpret_cast<uint32_t *>(smem + TILE * D);  // TILE*D is a multiple of 16
    __shared__ int s_stats[WARPS][NUM_STATS];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long tile0 = ((long long)blockIdx.x + p.tile_off) * TILE;
    const long long e = tile0 + tid;
    const bool valid = e < p.B;
    // Warps never wait for each other: each owns its 32 rows of the tile, its stat counters and
    // its own TMA stores, so a warp stalled on a deck byte does not hold back the other three.
    s_stats[warp][lane] = 0;
    // Programmatic dependent launch: this grid may start while the previous step's grid drains;
    // nothing it wrote may be read before griddepcontrol.wait (no-ops without the launch attribute).
    asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
    // L2 prefetch of the planes of the tile `pf_dist` CTAs ahead (about one wave of resident CTAs):
    // by the time that CTA is scheduled its 1 + N plane loads hit L2 instead of waiting for DRAM.
    // One bulk-prefetch instruction per plane (512 B = the tile's slice), issued by lane 0.
    if (p.pf_dist > 0 && lane == 0) {
        const long long pt = (long long)blockIdx.x + p.pf_dist;
        if (pt < (long long)gridDim.x) {
            const U128 *src = p.st.planes + (pt + p.tile_off) * TILE;
#pragma unroll
            for (int q = 0; q <= N; ++q)
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" ::"l"(src + (long long)q * p.Bpad),
                             "r"((uint32_t)(TILE * 16))
                             : "memory");
        }
    }
    // the policy's random word depends only on (seed, env, t): drawn before the wait, so CTAs that
    // were scheduled early have work to do while the previous step's grid drains
    // (t_base is only written between two replays of the graph this launch belongs to, never by a predecessor
    // that may still be running under programmatic dependent launch)
    uint32_t policy_rnd = 0u;
    if (POLICY) {
        unsigned long long t = p.t;
        if (p.t_base) t += *p.t_base;
        policy_rnd = policy_random(p.seed, p.first_env + (unsigned long long)e, t);
    }
    asm volatile("griddepcontrol.wait;\n" ::: "memory");

    Env<N> s;
    load_env_dev<N>(p.st.planes, p.Bpad, e, s);  // planes are padded to Bpad: in bounds for every thread
    int action = 0;
    if (!POLICY && valid) action = load_action(p.actions, p.action_dtype, e);
    __syncwarp();

    int act_class = -1;
    uint32_t dirty_rows = 0, pf_new = PF_KEEP;
    int done_code = SKYJO_RUNNING;
    constexpr bool ASSIST = N >= SKYJO_ASSIST_MIN_N;
    constexpr bool DEFER = !ASSIST && SKYJO_DEFER_SCORING != 0;
    Assist as;
    if (ASSIST) warp_assist<N, IND, POLICY>(p, s, e, valid, action, policy_rnd, lane, smem + (size_t)warp * 32 * D, as);
    Outcome oc;
    oc.scored = 0;
    if (valid) oc = env_step<N, IND, POLICY, ASSIST, DEFER>(p, e, s, action, policy_rnd, &as);
    if (DEFER) warp_score_deferred<N>(p, s, e, lane, smem + (size_t)warp * 32 * D, oc);
    if (valid) {
        dirty_rows = oc.dirty_rows;
        pf_new = oc.pf_new;
        done_code = oc.done_code;
        act_class = oc.act_class;
        if (oc.done_code != SKYJO_RUNNING || oc.reshuffled) {  // rare events
            int *st = s_stats[warp];
            if (oc.scored) {
                atomicAdd(&st[SKYJO_STAT_EPISODES], 1);
                atomicAdd(&st[SKYJO_STAT_EPISODE_STEPS], oc.ep_steps);
                atomicAdd(&st[SKYJO_STAT_SCORE_RAW_SUM], oc.raw_sum);
                atomicAdd(&st[SKYJO_STAT_WINNER_RAW_SUM], oc.winner_raw);
                atomicAdd(&st[SKYJO_STAT_FINISHER_RAW_SUM], oc.fin_raw);
                if (oc.penalised) {
                    atomicAdd(&st[SKYJO_STAT_PENALISED], 1);
                    atomicAdd(&st[SKYJO_STAT_PENALISED_RAW_SUM], oc.fin_raw);
                }
                atomicAdd(&st[SKYJO_STAT_REFUNDS], oc.refunds);
                if (oc.starter0) atomicAdd(&st[SKYJO_STAT_STARTER_SEAT0], 1);
                atomicAdd(&st[SKYJO_STAT_WINS_SEAT0 + oc.winner], 1);
            }
            if (oc.reshuffled) atomicAdd(&st[SKYJO_STAT_RESHUFFLES], 1);
            if (oc.done_code == SKYJO_DONE_ILLEGAL) atomicAdd(&st[SKYJO_STAT_ILLEGAL], 1);
            if (oc.done_code == SKYJO_DONE_TRUNCATED) atomicAdd(&st[SKYJO_STAT_TRUNCATED], 1);
        }
    }

    // ---- observation + mask of the next agent -------------------------------------------
    {
        OW ow;
        uint32_t next_hidden;
        encode_words<N, IND>(s, (int)(s.hdr >> HDR_CUR_SH) & 0xF, ow, &next_hidden);
        stage_stream<OW::NW, 3>(ow.s, s_obs, tid);
        stage_stream<7, 2>(ow.m, s_mask, tid);
        // The next agent is about to draw with no hidden card left: the next step ends this game
        // (skyjo.py:350-356) and installs the pre-dealt episode.  Pull its planes into L2 now.
        if (valid && install_next_step(p.auto_reset, s.hdr, next_hidden)) {
#pragma unroll
            for (int q = 0; q <= N; ++q)
                asm volatile("prefetch.global.L2 [%0];\n" ::"l"(p.st.next_planes + (long long)q * p.Bpad + e));
        }
    }
    // state write-back after the staging: a draw-pile prefetch issued by env_step has had the
    // whole encode to arrive before the header is stored
    if (valid) {
        store_env_dev<N>(p.st.planes, p.Bpad, e, s, dirty_rows, pf_new);
        p.agent[e] = (int8_t)((s.hdr >> HDR_CUR_SH) & 0xF);
        p.done[e] = (uint8_t)done_code;
    }
    // action-class counters: four byte lanes summed over the warp (REDUX)
    const unsigned cls = __reduce_add_sync(0xFFFFFFFFu, act_class >= 0 ? 1u << (8 * act_class) : 0u);
    fence_async_smem();
    __syncwarp();
    const long long w0 = tile0 + 32 * warp;
    const long long left = p.B - w0;
    const uint32_t rows_valid = left >= 32 ? 32u : (left > 0 ? (uint32_t)left : 0u);
    store_warp_slice(smem + (size_t)warp * 32 * D, p.obs + w0 * D, D, rows_valid, p.bulk_ok != 0, lane);
    store_warp_slice(smem + (size_t)TILE * D + (size_t)warp * 32 * 26, p.mask + w0 * 26, 26u, rows_valid,
                     p.bulk_ok != 0, lane);
    {
        long long v = s_stats[warp][lane];
        if (lane >= SKYJO_STAT_ACT_DRAW_PILE && lane <= SKYJO_STAT_ACT_FLIP)
            v = (cls >> (8 * (lane - SKYJO_STAT_ACT_DRAW_PILE))) & 0xFFu;
        if (lane == SKYJO_STAT_STEPS) v = (cls & 0xFFu) + ((cls >> 8) & 0xFFu) + ((cls >> 16) & 0xFFu) + (cls >> 24);
        if (v) atomicAdd(&p.st.stats[((blockIdx.x * WARPS + warp) % STAT_SLOTS) * NUM_STATS + lane], (unsigned long long)v);
    }
    if (lane == 0) {
        asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
    }
}

// ---- multi-step rollout with the in-kernel policy ------------------------------------------------
// K consecutive env-steps of the same 32 games per warp in ONE launch (the loop of
// sample_game.py:10-21 unrolled in time): the state is loaded once, stays in registers across the
// K steps and is written back once; every step still encodes the next agent's observation and
// action mask and stores them -- into slice k of time-major rollout tensors [K][B][...] -- exactly
// as K single-step launches would have published them one after the other.  With an in-kernel
// policy nothing outside the kernel consumes step k's outputs before step k+1, so the K - 1 state
// round trips through HBM, the launch ramps / tails and the exposed load latency between steps
// are pure overhead; this kernel removes them.  Rewards keep the [B, N] "last finished episode,
// cleared by the next step" semantics of the single-step kernel (skyjo_env.py:242-252).
template <int N, bool IND>
__global__ void __launch_bounds__(TILE, ROLLOUT_MIN_CTAS(N))
    rollout_kernel(const __grid_constant__ StepParams p, const __grid_constant__ RolloutParams r) {
    using OW = ObsWords<N, IND>;
    constexpr int D = OW::D;
    extern __shared__ __align__(128) uint8_t smem[];
    uint32_t *s_obs = reinterpret_cast<uint32_t *>(smem);
    uint32_t *s_mask = reinterpret_cast<uint32_t *>(smem + TILE * D);
    __shared__ int s_stats[WARPS][NUM_STATS];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long tile0 = (long long)blockIdx.x * TILE;
    const long long e = tile0 + tid;
    const bool valid = e < p.B;
    const unsigned long long genv = p.first_env + (unsigned long long)e;
    s_stats[warp][lane] = 0;
    asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
    if (p.pf_dist > 0 && lane == 0) {
        const long long pt = (long long)blockIdx.x + p.pf_dist;
        if (pt < (long long)gridDim.x) {
            const U128 *src = p.st.planes + pt * TILE;
#pragma unroll
            for (int q = 0; q <= N; ++q)
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" ::"l"(src + (long long)q * p.Bpad),
                             "r"((uint32_t)(TILE * 16))
                             : "memory");
        }
    }
    U4 pblk = policy_block(p.seed, genv, p.t);  // one Philox block per four lockstep steps
    uint32_t policy_rnd = policy_word(pblk, p.t);
    asm volatile("griddepcontrol.wait;\n" ::: "memory");

    Env<N> s;
    load_env_dev<N>(p.st.planes, p.Bpad, e, s);
    __syncwarp();

    const long long w0 = tile0 + 32 * warp;
    const long long left = p.B - w0;
    const uint32_t rows_valid = left >= 32 ? 32u : (left > 0 ? (uint32_t)left : 0u);
    uint32_t dirty_all = 0;
    long long lane_stat = 0;  // lane k accumulates statistics entry k over the K steps
    for (int k = 0; k < r.K; ++k) {
        if (k > 0) {
            const unsigned long long tk = p.t + (unsigned long long)k;
            if ((tk & 3ull) == 0ull) pblk = policy_block(p.seed, genv, tk);  // warp-uniform
            policy_rnd = policy_word(pblk, tk);
            // the staging tile is reused: the previous step's bulk stores must have read it
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
            __syncwarp();
        }
        int act_class = -1, done_code = SKYJO_RUNNING;
        uint32_t pf_new = PF_KEEP;
        constexpr bool ASSIST = N >= SKYJO_ASSIST_MIN_N;
        constexpr bool DEFER = !ASSIST && SKYJO_DEFER_SCORING != 0;
        Assist as;
        if (ASSIST) warp_assist<N, IND, true>(p, s, e, valid, 0, policy_rnd, lane, smem + (size_t)warp * 32 * D, as);
        Outcome oc;
        oc.scored = 0;
        if (valid) oc = env_step<N, IND, true, ASSIST, DEFER>(p, e, s, 0, policy_rnd, &as);
        if (DEFER) warp_score_deferred<N>(p, s, e, lane, smem + (size_t)warp * 32 * D, oc);
        if (valid) {
            dirty_all |= oc.dirty_rows;
            pf_new = oc.pf_new;
            done_code = oc.done_code;
            act_class = oc.act_class;
            if (oc.done_code != SKYJO_RUNNING || oc.reshuffled) {  // rare events
                int *st = s_stats[warp];
                if (oc.scored) {
                    atomicAdd(&st[SKYJO_STAT_EPISODES], 1);
                    atomicAdd(&st[SKYJO_STAT_EPISODE_STEPS], oc.ep_steps);
                    atomicAdd(&st[SKYJO_STAT_SCORE_RAW_SUM], oc.raw_sum);
                    atomicAdd(&st[SKYJO_STAT_WINNER_RAW_SUM], oc.winner_raw);
                    atomicAdd(&st[SKYJO_STAT_FINISHER_RAW_SUM], oc.fin_raw);
                    if (oc.penalised) {
                        atomicAdd(&st[SKYJO_STAT_PENALISED], 1);
                        atomicAdd(&st[SKYJO_STAT_PENALISED_RAW_SUM], oc.fin_raw);
                    }
                    atomicAdd(&st[SKYJO_STAT_REFUNDS], oc.refunds);
                    if (oc.starter0) atomicAdd(&st[SKYJO_STAT_STARTER_SEAT0], 1);
                    atomicAdd(&st[SKYJO_STAT_WINS_SEAT0 + oc.winner], 1);
                }
                if (oc.reshuffled) atomicAdd(&st[SKYJO_STAT_RESHUFFLES], 1);
                if (oc.done_code == SKYJO_DONE_ILLEGAL) atomicAdd(&st[SKYJO_STAT_ILLEGAL], 1);
                if (oc.done_code == SKYJO_DONE_TRUNCATED) atomicAdd(&st[SKYJO_STAT_TRUNCATED], 1);
            }
        }
        const bool last = k + 1 == r.K;
        {
            OW ow;
            uint32_t next_hidden;
            encode_words<N, IND>(s, (int)(s.hdr >> HDR_CUR_SH) & 0xF, ow, &next_hidden);
            stage_stream<OW::NW, 3>(ow.s, s_obs, tid);
            stage_stream<7, 2>(ow.m, s_mask, tid);
            if (valid && install_next_step(p.auto_reset, s.hdr, next_hidden)) {
#pragma unroll
                for (int q = 0; q <= N; ++q)
                    asm volatile("prefetch.global.L2 [%0];\n" ::"l"(p.st.next_planes + (long long)q * p.Bpad + e));
            }
        }
        // the byte below the drawn card (requested by env_step) becomes the prefetched pile top
        if (pf_new != PF_KEEP) s.hdr = hdr_pf_set(s.hdr, pf_new);
        if (valid) {
            const int8_t ag = (int8_t)((s.hdr >> HDR_CUR_SH) & 0xF);
            r.agent[(long long)k * p.B + e] = ag;
            r.done[(long long)k * p.B + e] = (uint8_t)done_code;
            if (last && r.publish) {
                p.agent[e] = ag;
                p.done[e] = (uint8_t)done_code;
            }
        }
        const unsigned cls = __reduce_add_sync(0xFFFFFFFFu, act_class >= 0 ? 1u << (8 * act_class) : 0u);
        if (lane >= SKYJO_STAT_ACT_DRAW_PILE && lane <= SKYJO_STAT_ACT_FLIP)
            lane_stat += (cls >> (8 * (lane - SKYJO_STAT_ACT_DRAW_PILE))) & 0xFFu;
        if (lane == SKYJO_STAT_STEPS)
            lane_stat += (cls & 0xFFu) + ((cls >> 8) & 0xFFu) + ((cls >> 16) & 0xFFu) + (cls >> 24);
        fence_async_smem();
        __syncwarp();
        store_warp_slice(smem + (size_t)warp * 32 * D, r.obs + ((long long)k * p.B + w0) * D, D, rows_valid,
                         r.bulk_ok != 0, lane);
        store_warp_slice(smem + (size_t)TILE * D + (size_t)warp * 32 * 26, r.mask + ((long long)k * p.B + w0) * 26, 26u,
                         rows_valid, r.bulk_ok != 0, lane);
        if (last && r.publish) {
            store_warp_slice(smem + (size_t)warp * 32 * D, p.obs + w0 * D, D, rows_valid, p.bulk_ok != 0, lane);
            store_warp_slice(smem + (size_t)TILE * D + (size_t)warp * 32 * 26, p.mask + w0 * 26, 26u, rows_valid,
                             p.bulk_ok != 0, lane);
        }
        if (lane == 0) asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
    }
    if (valid) store_env_dev<N>(p.st.planes, p.Bpad, e, s, dirty_all, PF_KEEP);
    {
        const long long v = lane_stat + s_stats[warp][lane];
        if (v) atomicAdd(&p.st.stats[((blockIdx.x * WARPS + warp) % STAT_SLOTS) * NUM_STATS + lane], (unsigned long long)v);
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
}

// Stand-alone observe (SimpleSkyjoEnv.observe, skyjo_env.py:199-214): encodes the view of
// `agent` (or of each env's agent_selection when agent < 0) without touching the state.
// With reset_outputs it also publishes agent / done = 0 / reward = 0 (after a reset).
template <int N, bool IND>
__global__ void __launch_bounds__(TILE) observe_kernel(const StepParams p, int agent, int8_t *obs_out, int8_t *mask_out,
                                                       int reset_outputs, int bulk_ok) {
    using OW = ObsWords<N, IND>;
    constexpr int D = OW::D;
    extern __shared__ __align__(128) uint8_t smem[];
    uint32_t *s_obs = reinterpret_cast<uint32_t *>(smem);
    uint32_t *s_mask = reinterpret_cast<uint32_t *>(smem + TILE * D);
    const int tid = threadIdx.x;
    const long long tile0 = (long long)blockIdx.x * TILE;
    const long long e = tile0 + tid;
    Env<N> s;
    load_env<N>(p.st.planes, p.Bpad, e, s);
    const int cur = (int)(s.hdr >> HDR_CUR_SH) & 0xF;
    if (reset_outputs && e < p.B) {
        p.agent[e] = (int8_t)cur;
        p.done[e] = 0;
#pragma unroll
        for (int q = 0; q < N; ++q) {
            p.reward[e * N + q] = 0.0;
            p.final_score[e * N + q] = 0.0;
        }
    }
    {
        OW ow;
        encode_words<N, IND>(s, agent < 0 ? cur : agent, ow);
        stage_stream<OW::NW, 3>(ow.s, s_obs, tid);
        stage_stream<7, 2>(ow.m, s_mask, tid);
    }
    fence_async_smem();
    __syncthreads();
    const long long left = p.B - tile0;
    const uint32_t n_env = left >= TILE ? TILE : (left > 0 ? (uint32_t)left : 0u);  // CTAs may lie in the padding
    const bool bulk = bulk_ok && n_env == TILE;
    store_tile(smem, obs_out + tile0 * D, n_env * D, bulk, tid);
    store_tile(smem + TILE * D, mask_out + tile0 * 26, n_env * 26u, bulk, tid);
    if (bulk) bulk_commit_and_wait(tid);
}

// launchers implemented per player count in skyjo_step_inst.cu
typedef cudaError_t (*step_launch_fn)(const StepParams &, bool indirect, bool policy, cudaStream_t);
typedef cudaError_t (*rollout_launch_fn)(const StepParams &, const RolloutParams &, bool indirect, cudaStream_t);
typedef cudaError_t (*observe_launch_fn)(const StepParams &, bool indirect, int agent, int8_t *, int8_t *, int, int,
                                         cudaStream_t);

}  // namespace skyjo
