// tests/hostsim/hostsim.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Compiles the per-env device functions of the CUDA library (skyjo_rl_b200/csrc/skyjo_core.cuh,
// skyjo_deal.cuh: they are __host__ __device__) with g++ and runs them in plain loops over a
// host copy of the HBM layout, including an emulation of the warp-shuffle staging of the
// observation tiles.  It exists so that `pytest -m "not gpu"` can check the kernel logic
// against the oracle in a container without a GPU.  The product never loads this library:
// skyjo_rl_b200 only binds libskyjo_b200.so, whose compute entries fail without a CUDA device.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../include/skyjo_b200.h"
#include "../../skyjo_rl_b200/csrc/skyjo_core.cuh"
#include "../../skyjo_rl_b200/csrc/skyjo_deal.cuh"

using namespace skyjo;

struct HostSim {
    SkyjoConfig cfg;
    long long B, Bpad;
    unsigned long long seed, first_env, t;
    int D;
    std::vector<U128> planes, next_planes;
    std::vector<uint8_t> deck, needs_deal;
    std::vector<uint32_t> episode;
    uint32_t errflag;
    std::vector<int8_t> obs, mask, agent;
    std::vector<uint8_t> done;
    std::vector<double> reward, score;
    long long stats[NUM_STATS];
    DeviceState st;
};

struct ArrayDeck {
    uint8_t b[152];
    uint8_t get(int i) const { return b[i]; }
    void set(int i, uint8_t v) { b[i] = v; }
    uint8_t get_at(int ibase, int off) const { return b[ibase + off]; }
    void set_at(int ibase, int off, uint8_t v) { b[ibase + off] = v; }
    uint32_t word(int w) const {
        uint32_t x;
        memcpy(&x, b + 4 * w, 4);
        return x;
    }
};

static StepParams make_params(HostSim *h) {
    StepParams p;
    memset(&p, 0, sizeof(p));
    p.st = h->st;
    p.obs = h->obs.data();
    p.mask = h->mask.data();
    p.agent = h->agent.data();
    p.done = h->done.data();
    p.reward = h->reward.data();
    p.final_score = h->score.data();
    p.B = h->B;
    p.Bpad = h->Bpad;
    p.first_env = h->first_env;
    p.seed = h->seed;
    p.t = h->t;
    p.score_penalty = h->cfg.score_penalty;
    p.mean_reward = h->cfg.mean_reward;
    p.reward_refunded = h->cfg.reward_refunded;
    p.auto_reset = h->cfg.auto_reset;
    p.max_steps = h->cfg.max_episode_steps;
    return p;
}

static void deal(HostSim *h, int flagged, int target_next, const int8_t *decks, const uint8_t *flips) {
    DealParams d;
    d.st = h->st;
    d.B = h->B;
    d.Bpad = h->Bpad;
    d.first_env = h->first_env;
    d.seed = h->seed;
    d.N = h->cfg.num_players;
    d.indirect = h->cfg.observe_other_player_indirect ? 1 : 0;
    d.flagged = flagged;
    d.target_next = target_next;
    d.decks = decks;
    d.flips = flips;
    for (long long e = 0; e < h->B; ++e) {
        ArrayDeck deck;
        if (!flagged) {
            deal_one(d, e, target_next ? 1u : 0u, deck);
        } else if (h->needs_deal[e]) {
            deal_one(d, e, (uint32_t)(h->needs_deal[e] >> 1), deck);
            h->needs_deal[e] = 0;
        }
    }
}

// emulation of stage_stream + the tile store for one tensor
template <int NW, int TAIL>
static void stage_tile(const std::vector<uint32_t> &S /* [rows][NW] */, int rows_valid, int8_t *dst) {
    constexpr int D = 4 * NW - 4 + TAIL;
    std::vector<uint32_t> tile(TILE * D / 4 + 4, 0xDEADBEEFu);
    for (int tid = 0; tid < TILE; ++tid) {
        uint32_t own[NW], out[NW], first;
        int count;
        for (int k = 0; k < NW; ++k) own[k] = S[(size_t)tid * NW + k];
        const int lane = tid & 31;
        const uint32_t next0 = lane < 31 ? S[(size_t)(tid + 1) * NW] : own[0];  // __shfl_down_sync semantics
        stage_words<NW, TAIL>(own, next0, tid, out, first, count);
        for (int k = 0; k < count; ++k) tile[first + k] = out[k];
    }
    memcpy(dst, tile.data(), (size_t)rows_valid * D);
}

template <int N, bool IND>
static void encode_all(HostSim *h, int agent, int8_t *obs_out, int8_t *mask_out) {
    using OW = ObsWords<N, IND>;
    for (long long tile0 = 0; tile0 < h->B; tile0 += TILE) {
        std::vector<uint32_t> S((size_t)TILE * OW::NW), M((size_t)TILE * 7);
        for (int tid = 0; tid < TILE; ++tid) {
            Env<N> s;
            load_env<N>(h->st.planes, h->Bpad, tile0 + tid, s);
            OW ow;
            const int cur = (int)(s.hdr >> HDR_CUR_SH) & 0xF;
            encode_words<N, IND>(s, agent < 0 ? cur : agent, ow);
            memcpy(&S[(size_t)tid * OW::NW], ow.s, sizeof(ow.s));
            memcpy(&M[(size_t)tid * 7], ow.m, sizeof(ow.m));
        }
        const long long left = h->B - tile0;
        const int n_env = left >= TILE ? TILE : (int)left;
        stage_tile<OW::NW, 3>(S, n_env, obs_out + tile0 * OW::D);
        stage_tile<7, 2>(M, n_env, mask_out + tile0 * 26);
    }
}

static void add_stats(HostSim *h, const Outcome &oc) {
    long long *st = h->stats;
    if (oc.act_class >= 0) {
        st[SKYJO_STAT_ACT_DRAW_PILE + oc.act_class] += 1;
        st[SKYJO_STAT_STEPS] += 1;
    }
    if (oc.scored) {
        st[SKYJO_STAT_EPISODES] += 1;
        st[SKYJO_STAT_EPISODE_STEPS] += oc.ep_steps;
        st[SKYJO_STAT_SCORE_RAW_SUM] += oc.raw_sum;
        st[SKYJO_STAT_WINNER_RAW_SUM] += oc.winner_raw;
        st[SKYJO_STAT_FINISHER_RAW_SUM] += oc.fin_raw;
        if (oc.penalised) {
            st[SKYJO_STAT_PENALISED] += 1;
            st[SKYJO_STAT_PENALISED_RAW_SUM] += oc.fin_raw;
        }
        st[SKYJO_STAT_REFUNDS] += oc.refunds;
        if (oc.starter0) st[SKYJO_STAT_STARTER_SEAT0] += 1;
        st[SKYJO_STAT_WINS_SEAT0 + oc.winner] += 1;
    }
    if (oc.reshuffled) st[SKYJO_STAT_RESHUFFLES] += 1;
    if (oc.done_code == SKYJO_DONE_ILLEGAL) st[SKYJO_STAT_ILLEGAL] += 1;
    if (oc.done_code == SKYJO_DONE_TRUNCATED) st[SKYJO_STAT_TRUNCATED] += 1;
}

template <int N, bool IND, bool POLICY>
static void step_all(HostSim *h, const int32_t *actions) {
    StepParams p = make_params(h);
    p.actions = actions;
    p.action_dtype = SKYJO_ACT_I32;
    for (long long e = 0; e < h->B; ++e) {
        Env<N> s;
        load_env<N>(h->st.planes, h->Bpad, e, s);
        const int action = POLICY ? 0 : load_action(actions, SKYJO_ACT_I32, e);
        const Outcome oc = env_step<N, IND, POLICY>(p, e, s, action,
                                                    POLICY ? policy_random(p.seed, p.first_env + (unsigned long long)e, p.t) : 0u);
        store_env<N>(h->st.planes, h->Bpad, e, s, oc.dirty_rows, oc.pf_new);
        h->agent[e] = (int8_t)((s.hdr >> HDR_CUR_SH) & 0xF);
        h->done[e] = (uint8_t)oc.done_code;
        add_stats(h, oc);
    }
    encode_all<N, IND>(h, -1, h->obs.data(), h->mask.data());
    h->t += 1;
    if (h->cfg.auto_reset) deal(h, 1, 1, nullptr, nullptr);
}

#define HS_DISPATCH_N(FN, ...)                      \
    switch (h->cfg.num_players) {                   \
        case 1: FN(1, __VA_ARGS__); break;          \
        case 2: FN(2, __VA_ARGS__); break;          \
        case 3: FN(3, __VA_ARGS__); break;          \
        case 4: FN(4, __VA_ARGS__); break;          \
        case 5: FN(5, __VA_ARGS__); break;          \
        case 6: FN(6, __VA_ARGS__); break;          \
        case 7: FN(7, __VA_ARGS__); break;          \
        case 8: FN(8, __VA_ARGS__); break;          \
        case 9: FN(9, __VA_ARGS__); break;          \
        case 10: FN(10, __VA_ARGS__); break;        \
        case 11: FN(11, __VA_ARGS__); break;        \
        default: FN(12, __VA_ARGS__); break;        \
    }

#define HS_STEP(NN, ACT)                                                  \
    do {                                                                  \
        if (ind) {                                                        \
            if (ACT) step_all<NN, true, false>(h, ACT);                   \
            else step_all<NN, true, true>(h, nullptr);                    \
        } else {                                                          \
            if (ACT) step_all<NN, false, false>(h, ACT);                  \
            else step_all<NN, false, true>(h, nullptr);                   \
        }                                                                 \
    } while (0)

#define HS_OBSERVE(NN, AG, O, M)                                          \
    do {                                                                  \
        if (ind) encode_all<NN, true>(h, AG, O, M);                       \
        else encode_all<NN, false>(h, AG, O, M);                          \
    } while (0)

extern "C" {

HostSim *hs_create(const SkyjoConfig *cfg, long long B, unsigned long long seed, long long first_env) {
    HostSim *h = new HostSim();
    h->cfg = *cfg;
    h->B = B;
    h->Bpad = (B + ENV_PAD - 1) / ENV_PAD * ENV_PAD;
    h->seed = seed;
    h->first_env = (unsigned long long)first_env;
    h->t = 0;
    const int N = cfg->num_players;
    h->D = cfg->observe_other_player_indirect ? 31 : 19 + 12 * N;
    h->planes.assign((size_t)num_planes(N) * h->Bpad, U128{0, 0, 0, 0});
    h->next_planes.assign((size_t)num_planes(N) * h->Bpad, U128{0, 0, 0, 0});
    h->deck.assign((size_t)2 * h->Bpad * PILE_ROW + 16, 0);
    h->needs_deal.assign((size_t)h->Bpad, 0);
    h->episode.assign((size_t)h->Bpad, 0);
    h->errflag = 0;
    h->obs.assign((size_t)B * h->D, 0);
    h->mask.assign((size_t)B * 26, 0);
    h->agent.assign((size_t)B, 0);
    h->done.assign((size_t)B, 0);
    h->reward.assign((size_t)B * N, 0.0);
    h->score.assign((size_t)B * N, 0.0);
    memset(h->stats, 0, sizeof(h->stats));
    h->st.planes = h->planes.data();
    h->st.next_planes = h->next_planes.data();
    // 16-byte aligned deck base (rows are written with 16-byte stores)
    uintptr_t base = (uintptr_t)h->deck.data();
    h->st.deck = (uint8_t *)((base + 15) & ~(uintptr_t)15);
    h->st.episode = h->episode.data();
    h->st.needs_deal = h->needs_deal.data();
    h->st.stats = nullptr;
    h->st.errflag = &h->errflag;
    return h;
}

void hs_destroy(HostSim *h) { delete h; }

static void after_reset(HostSim *h) {
    const bool ind = h->cfg.observe_other_player_indirect != 0;
    if (h->cfg.auto_reset) deal(h, 0, 1, nullptr, nullptr);
    for (long long e = 0; e < h->B; ++e) {
        const U128 P0 = h->planes[(size_t)e];
        h->agent[e] = (int8_t)((P0.x >> HDR_CUR_SH) & 0xF);
        h->done[e] = 0;
    }
    std::fill(h->reward.begin(), h->reward.end(), 0.0);
    std::fill(h->score.begin(), h->score.end(), 0.0);
    HS_DISPATCH_N(HS_OBSERVE, -1, h->obs.data(), h->mask.data());
}

void hs_reset(HostSim *h) {
    std::fill(h->needs_deal.begin(), h->needs_deal.end(), 0);
    deal(h, 0, 0, nullptr, nullptr);
    after_reset(h);
}

void hs_reset_injected(HostSim *h, const int8_t *decks, const uint8_t *flips) {
    std::fill(h->needs_deal.begin(), h->needs_deal.end(), 0);
    deal(h, 0, 0, decks, flips);
    after_reset(h);
}

// actions == NULL: uniform legal policy drawn from the product RNG (skyjo_step_random)
void hs_step(HostSim *h, const int32_t *actions) {
    const bool ind = h->cfg.observe_other_player_indirect != 0;
    HS_DISPATCH_N(HS_STEP, actions);
}

void hs_observe(HostSim *h, int agent, int8_t *obs_out, int8_t *mask_out) {
    const bool ind = h->cfg.observe_other_player_indirect != 0;
    HS_DISPATCH_N(HS_OBSERVE, agent, obs_out, mask_out);
}

void hs_export(HostSim *h, long long e, SkyjoEnvDebug *out) {
    export_one(h->st, h->Bpad, h->cfg.num_players, h->cfg.observe_other_player_indirect ? 1 : 0, e, *out);
}

int8_t *hs_obs(HostSim *h) { return h->obs.data(); }
int8_t *hs_mask(HostSim *h) { return h->mask.data(); }
int8_t *hs_agent(HostSim *h) { return h->agent.data(); }
uint8_t *hs_done(HostSim *h) { return h->done.data(); }
double *hs_reward(HostSim *h) { return h->reward.data(); }
double *hs_score(HostSim *h) { return h->score.data(); }
long long *hs_stats(HostSim *h) { return h->stats; }
uint32_t hs_errflag(HostSim *h) { return h->errflag; }
int hs_obs_len(HostSim *h) { return h->D; }

}  // extern "C"
