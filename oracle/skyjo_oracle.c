/*
 * oracle/skyjo_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see skyjo_oracle.h).
 *
 * Single-game CPU restatement of rlskyjo.game.skyjo.SkyjoGame.  The piles are kept as
 * real stacks exactly like the reference's Python lists, masks and cards as (N,12) int8
 * arrays; nothing here shares a data layout or a line of code with the CUDA kernels.
 * Citations are into /root/reference.
 */
#include "skyjo_oracle.h"

#include <string.h>

/* ------------------------------------------------------------------------------------ */
int sk_obs_len(int num_players, int indirect) {
    /* skyjo.py:43-45 */
    return indirect ? 19 + 12 : 19 + 12 * num_players;
}

int sk_init(sk_game *g, int num_players, double score_penalty, int indirect) {
    /* skyjo.py:24-26 */
    if (!(0 < num_players && num_players <= SK_MAX_PLAYERS)) return SK_ERR_RANGE;
    memset(g, 0, sizeof(*g));
    g->num_players = num_players;
    g->score_penalty = score_penalty;
    g->indirect = indirect ? 1 : 0;
    g->hand = SK_UNK;
    return SK_OK;
}

/* skyjo.py:226-257 _jit_observe_global_game_stats */
static void observe_global_game_stats(const sk_game *g, int count_players_cards, int counts[15],
                                      long known_sum[SK_MAX_PLAYERS],
                                      long n_hidden[SK_MAX_PLAYERS], int *pile_top) {
    /* :236 counted = pile + range(-2, 13); :248 bincount(counted - min) - 1 */
    int bins[15];
    for (int j = 0; j < 15; ++j) bins[j] = 1;
    for (int i = 0; i < g->n_disc; ++i) bins[g->discard[i] + 2] += 1;
    for (int pl = 0; pl < g->num_players; ++pl) {
        long s = 0;
        for (int i = 0; i < SK_SLOTS; ++i) {
            if (g->masked[pl][i] == 1) { /* :240 masked_option */
                if (count_players_cards) bins[g->cards[pl][i] + 2] += 1; /* :243-244 */
                s += g->cards[pl][i];                                      /* :246 */
            }
        }
        known_sum[pl] = s;
    }
    for (int j = 0; j < 15; ++j) counts[j] = bins[j] - 1;
    for (int pl = 0; pl < g->num_players; ++pl) { /* :250-252 */
        long h = 0;
        for (int i = 0; i < SK_SLOTS; ++i) h += (g->masked[pl][i] == 2);
        n_hidden[pl] = h;
    }
    *pile_top = g->n_disc ? g->discard[g->n_disc - 1] : SK_EMPTY_TOP; /* :254 */
}

/* skyjo.py:127-138 _reshuffle_discard_pile: shuffle, drawpile=list(pile), discard=[pop()] */
static void reshuffle_into_drawpile(sk_game *g, int8_t *pile, int len) {
    if (g->shuffle) g->shuffle(g->shuffle_ctx, pile, len); /* :135 */
    memcpy(g->drawpile, pile, (size_t)len);                /* :136 */
    g->n_draw = len - 1;                                   /* :137 pop() */
    g->discard[0] = pile[len - 1];
    g->n_disc = 1;
}

/* skyjo.py:105-125 _reset_start_player */
static void reset_start_player(sk_game *g) {
    int counts[15], top;
    long ksum[SK_MAX_PLAYERS], nhid[SK_MAX_PLAYERS];
    observe_global_game_stats(g, 1, counts, ksum, nhid, &top);
    int arg = 0; /* :112 argmax, first maximum wins */
    for (int pl = 1; pl < g->num_players; ++pl)
        if (ksum[pl] > ksum[arg]) arg = pl;
    /* :114-125: cycle [0,draw],[0,place],[1,draw].. advanced 1 + 2*arg times */
    g->exp_player = arg;
    g->exp_phase = SK_PHASE_DRAW;
}

int sk_reset_injected(sk_game *g, const int8_t *deck150, const uint8_t *flips) {
    const int N = g->num_players;
    /* skyjo.py:54-61 */
    g->is_terminated = 0;
    g->has_final_score = 0;
    g->n_reshuffles = 0;
    for (int p = 0; p < SK_MAX_PLAYERS; ++p) {
        g->num_refunded[p] = 0;
        g->num_placed[p] = 0;
        g->final_score[p] = 0.0;
    }
    g->hand = SK_UNK;
    /* :63-65 players_cards = deck[:12N].reshape(N, 12) */
    for (int p = 0; p < N; ++p)
        for (int i = 0; i < SK_SLOTS; ++i) g->cards[p][i] = deck150[12 * p + i];
    /* :68-70 with the identity permutation in place of the second shuffle */
    int8_t rest[SK_DECK];
    int len = SK_DECK - 12 * N;
    memcpy(rest, deck150 + 12 * N, (size_t)len);
    sk_shuffle_fn keep = g->shuffle;
    g->shuffle = 0;
    reshuffle_into_drawpile(g, rest, len);
    g->shuffle = keep;
    /* :96-103 _reset_card_mask: all 2, two picked slots 1 */
    for (int p = 0; p < N; ++p) {
        for (int i = 0; i < SK_SLOTS; ++i) g->masked[p][i] = 2;
        uint8_t a = flips[2 * p], b = flips[2 * p + 1];
        if (a >= SK_SLOTS || b >= SK_SLOTS || a == b) return SK_ERR_RANGE;
        g->masked[p][a] = 1;
        g->masked[p][b] = 1;
    }
    reset_start_player(g); /* :73 */
    return SK_OK;
}

int sk_collect_observation(const sk_game *g, int player_id, int8_t *obs, int8_t *mask) {
    if (player_id < 0 || player_id >= g->num_players) return SK_ERR_RANGE;
    int counts[15], top;
    long ksum[SK_MAX_PLAYERS], nhid[SK_MAX_PLAYERS];
    observe_global_game_stats(g, !g->indirect, counts, ksum, nhid, &top); /* :151-161 */
    long smin = ksum[0], hmin = nhid[0];
    for (int p = 1; p < g->num_players; ++p) {
        if (ksum[p] < smin) smin = ksum[p];
        if (nhid[p] < hmin) hmin = nhid[p];
    }
    int k = 0;
    obs[k++] = (int8_t)(smin < 127 ? smin : 127);           /* :182 */
    obs[k++] = (int8_t)hmin;                                 /* :183 */
    for (int j = 0; j < 15; ++j) obs[k++] = (int8_t)counts[j]; /* :184 */
    obs[k++] = (int8_t)top;                                  /* :185 */
    obs[k++] = (int8_t)g->hand;                              /* :186 */
    if (g->indirect) {
        /* :259-277 own row only */
        for (int i = 0; i < SK_SLOTS; ++i)
            obs[k++] = g->masked[player_id][i] != 2 ? g->cards[player_id][i] : (int8_t)SK_UNK;
    } else {
        /* :279-302 all rows; the np.roll only permutes the iteration order, the output
         * stays in absolute player order */
        for (int p = 0; p < g->num_players; ++p)
            for (int i = 0; i < SK_SLOTS; ++i)
                obs[k++] = g->masked[p][i] != 2 ? g->cards[p][i] : (int8_t)SK_UNK;
    }
    /* :201-224 _jit_action_mask */
    if (g->exp_phase == SK_PHASE_PLACE) {
        for (int i = 0; i < SK_SLOTS; ++i) {
            mask[i] = g->masked[player_id][i] != 0;
            mask[12 + i] = g->masked[player_id][i] == 2;
        }
        mask[24] = 0;
        mask[25] = 0;
    } else {
        memset(mask, 0, 24);
        mask[24] = 1;
        mask[25] = 1;
    }
    return SK_OK;
}

void sk_evaluate_game(const int8_t *cards, int num_players, int finisher, double score_penalty,
                      double *score) {
    /* skyjo.py:486-493 */
    for (int pl = 0; pl < num_players; ++pl) {
        score[pl] = 0.0;
        for (int st = 0; st < 4; ++st) {
            const int8_t *c = cards + 12 * pl + 3 * st;
            int mn = c[0], mx = c[0];
            for (int i = 1; i < 3; ++i) {
                if (c[i] < mn) mn = c[i];
                if (c[i] > mx) mx = c[i];
            }
            if (mn != mx) score[pl] += (double)(c[0] + c[1] + c[2]);
        }
    }
    /* :496-497 */
    double mn = score[0];
    for (int pl = 1; pl < num_players; ++pl)
        if (score[pl] < mn) mn = score[pl];
    if (mn != score[finisher]) score[finisher] *= score_penalty;
}

/* skyjo.py:471-475 */
static int player_goal_check(const sk_game *g, int player_id) {
    for (int i = 0; i < SK_SLOTS; ++i)
        if (g->masked[player_id][i] == 2) return 0;
    return 1;
}

/* skyjo.py:142-144 + the cycle built at :114-120 */
static void internal_next_action(sk_game *g) {
    if (g->exp_phase == SK_PHASE_DRAW) {
        g->exp_phase = SK_PHASE_PLACE;
    } else {
        g->exp_phase = SK_PHASE_DRAW;
        g->exp_player = (g->exp_player + 1) % g->num_players;
    }
}

/* skyjo.py:337-374 */
static int action_draw_card(sk_game *g, int player_id, int draw_from) {
    if (player_goal_check(g, player_id)) { /* :350-356 */
        g->is_terminated = 1;
        sk_evaluate_game(&g->cards[0][0], g->num_players, player_id, g->score_penalty,
                         g->final_score);
        g->has_final_score = 1;
        return SK_GAME_OVER;
    }
    if (draw_from == 24) {
        if (g->n_draw == 0) { /* :361-365 */
            int8_t tmp[SK_PILE_CAP];
            int len = g->n_disc;
            memcpy(tmp, g->discard, (size_t)len);
            reshuffle_into_drawpile(g, tmp, len);
            g->n_reshuffles += 1;
        }
        g->hand = g->drawpile[--g->n_draw]; /* :366 */
    } else {
        g->hand = g->discard[--g->n_disc]; /* :370 */
    }
    internal_next_action(g); /* :373 */
    return SK_OK;
}

/* skyjo.py:431-469 */
static int remask_refunded(sk_game *g, int player_id, int8_t *to_discard, int *n_to_discard) {
    int updated = 0;
    *n_to_discard = 0;
    for (int st = 0; st < 4; ++st) {
        int8_t *c = &g->cards[player_id][3 * st];
        int8_t *m = &g->masked[player_id][3 * st];
        int mn = c[0], mx = c[0];
        for (int i = 1; i < 3; ++i) {
            if (c[i] < mn) mn = c[i];
            if (c[i] > mx) mx = c[i];
        }
        if (mn == mx && m[0] == 1 && m[1] == 1 && m[2] == 1) { /* :451-453 */
            for (int i = 0; i < 3; ++i) m[i] = 0;              /* :454 */
            /* :456-458 appends the already-zeroed MASK slice, i.e. three 0s */
            for (int i = 0; i < 3; ++i) to_discard[(*n_to_discard)++] = m[i];
            for (int i = 0; i < 3; ++i) c[i] = SK_REFUNDED;    /* :459 */
            updated = 1;
        }
    }
    return updated;
}

/* skyjo.py:376-427 */
static int action_place(sk_game *g, int player_id, int a) {
    if (a < 12) { /* :389-395 */
        g->discard[g->n_disc++] = g->cards[player_id][a];
        g->masked[player_id][a] = 1;
        g->cards[player_id][a] = (int8_t)g->hand;
    } else { /* :396-404 */
        int pos = a - 12;
        if (g->masked[player_id][pos] != 2) return SK_ERR_REVEALED; /* :399 */
        g->discard[g->n_disc++] = (int8_t)g->hand;
        g->masked[player_id][pos] = 1;
    }
    int8_t add[12];
    int n_add;
    if (remask_refunded(g, player_id, add, &n_add)) { /* :407-421 */
        g->num_refunded[player_id] += 1;
        for (int i = 0; i < n_add; ++i) g->discard[g->n_disc++] = add[i];
    }
    g->num_placed[player_id] += 1; /* :424 */
    g->hand = SK_UNK;              /* :425 */
    internal_next_action(g);       /* :426 */
    return SK_OK;
}

int sk_act(sk_game *g, int player_id, int action) {
    if (g->exp_player != player_id) return SK_ERR_TURN;  /* :310-313 */
    if (action < 0 || action > 25) return SK_ERR_RANGE;  /* :314 */
    if (g->is_terminated) return SK_GAME_OVER;           /* :316-321 (warns) */
    if (action >= 24) {
        if (g->hand != SK_UNK) return SK_ERR_PHASE;      /* :324-328 */
        return action_draw_card(g, player_id, action);
    }
    if (g->hand == SK_UNK) return SK_ERR_PHASE;          /* :331-334 */
    return action_place(g, player_id, action);
}

/* numpy's add.reduce for a contiguous float64 vector of n <= 128 elements:
 * plain left-to-right below 8, eight accumulators + tail from 8 (checked against
 * numpy 2.3.5 in tests/test_oracle_golden.py::test_numpy_sum_order). */
static double np_sum(const double *a, int n) {
    if (n < 8) {
        double res = 0.0;
        for (int i = 0; i < n; ++i) res += a[i];
        return res;
    }
    double r[8];
    for (int k = 0; k < 8; ++k) r[k] = a[k];
    int i = 8;
    for (; i < n - (n % 8); i += 8)
        for (int k = 0; k < 8; ++k) r[k] += a[i + k];
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += a[i];
    return res;
}

void sk_calc_final_rewards(const double *final_score, const int *num_refunded, int num_players,
                           double mean_reward, double reward_refunded, double *reward) {
    /* skyjo_env.py:307-308: reward = -score + np.mean(score) + self.mean_reward */
    double mean = np_sum(final_score, num_players) / (double)num_players;
    for (int p = 0; p < num_players; ++p) reward[p] = (-final_score[p] + mean) + mean_reward;
    /* :310-311 */
    if (reward_refunded != 0.0)
        for (int p = 0; p < num_players; ++p)
            reward[p] += (double)num_refunded[p] * reward_refunded;
}

/* ------------------------------------------------------------------------------------
 * Twins of the product RNG.  The reference shuffles with numba's global MT19937
 * (skyjo.py:81,94,101,135), which north_star replaces by counter-based Philox; these
 * functions restate the PRODUCT's deal so that the oracle can be driven on identical
 * decks.  Independent restatement of csrc/skyjo_rng.cuh (do not share code with it).
 * Philox4x32-10: Salmon et al., "Parallel random numbers: as easy as 1, 2, 3" (SC11).
 * ------------------------------------------------------------------------------------ */
void sk_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

enum { PURPOSE_DEAL = 1, PURPOSE_FLIPS = 2, PURPOSE_RESHUFFLE = 3, PURPOSE_POLICY = 4 };

static void rng_block(uint64_t seed, uint64_t env, int purpose, uint32_t a, uint32_t b,
                      uint32_t out[4]) {
    uint32_t ctr[4] = {(uint32_t)env, (uint32_t)((env >> 32) & 0xFFFFFFu) | ((uint32_t)purpose << 24),
                       a, b};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    sk_philox4x32_10(ctr, key, out);
}

static uint32_t bounded(uint32_t r, uint32_t n) { return (uint32_t)(((uint64_t)r * n) >> 32); }

void sk_rng_deck(uint64_t seed, uint64_t env, uint32_t episode, int8_t *deck) {
    /* skyjo.py:80: ten each of -2..12, then a Fisher-Yates pass from the top */
    for (int i = 0; i < SK_DECK; ++i) deck[i] = (int8_t)(i / 10 - 2);
    uint32_t blk[4];
    for (int i = SK_DECK - 1; i >= 1; --i) {
        int k = SK_DECK - 1 - i;
        if ((k & 3) == 0) rng_block(seed, env, PURPOSE_DEAL, episode, (uint32_t)(k >> 2), blk);
        uint32_t j = bounded(blk[k & 3], (uint32_t)(i + 1));
        int8_t t = deck[i];
        deck[i] = deck[j];
        deck[j] = t;
    }
}

void sk_rng_flips(uint64_t seed, uint64_t env, uint32_t episode, int num_players, uint8_t *flips) {
    /* skyjo.py:101 choice(12, 2, replace=False) */
    for (int p = 0; p < num_players; ++p) {
        uint32_t blk[4];
        rng_block(seed, env, PURPOSE_FLIPS, episode, (uint32_t)p, blk);
        uint32_t a = bounded(blk[0], 12), b = bounded(blk[1], 11);
        if (b >= a) b += 1;
        flips[2 * p] = (uint8_t)a;
        flips[2 * p + 1] = (uint8_t)b;
    }
}

int sk_rng_policy(uint64_t seed, uint64_t env, uint64_t t, const int8_t *mask26) {
    /* random_admissible_policy.py:26-28: uniform over the legal actions */
    int count = 0;
    for (int i = 0; i < 26; ++i) count += mask26[i] != 0;
    if (count == 0) return -1;
    uint32_t blk[4];
    /* product protocol: one Philox block per four lockstep steps, step t takes word t & 3 */
    rng_block(seed, env, PURPOSE_POLICY, (uint32_t)(t >> 2), (uint32_t)(t >> 34), blk);
    int k = (int)bounded(blk[t & 3u], (uint32_t)count);
    for (int i = 0; i < 26; ++i)
        if (mask26[i] && k-- == 0) return i;
    return -1;
}

void sk_rng_reshuffle(void *vctx, int8_t *pile, int len) {
    /* Product rule for skyjo.py:135: the new order is drawn by sampling the pile's
     * multiset without replacement; e_0 becomes the new discard card, e_1.. are the
     * successive draws, so the python list is [e_{len-1}, ..., e_1, e_0]. */
    sk_rng_shuffle_ctx *ctx = (sk_rng_shuffle_ctx *)vctx;
    int bins[15] = {0};
    for (int i = 0; i < len; ++i) bins[pile[i] + 2] += 1;
    uint32_t q = ctx->q & 0x7Fu; /* the product keeps 7 bits of the reshuffle index */
    for (int d = 0; d < len; ++d) {
        uint32_t remaining = (uint32_t)(len - d);
        uint32_t blk[4];
        rng_block(ctx->seed, ctx->env, PURPOSE_RESHUFFLE, ctx->episode, (q << 16) | remaining, blk);
        uint32_t idx = bounded(blk[0], remaining);
        int j = 0;
        while (idx >= (uint32_t)bins[j]) {
            idx -= (uint32_t)bins[j];
            ++j;
        }
        bins[j] -= 1;
        pile[len - 1 - d] = (int8_t)(j - 2);
    }
    ctx->q += 1;
}

int sk_reset_rng(sk_game *g, sk_rng_shuffle_ctx *ctx) {
    int8_t deck[SK_DECK];
    uint8_t flips[2 * SK_MAX_PLAYERS];
    sk_rng_deck(ctx->seed, ctx->env, ctx->episode, deck);
    sk_rng_flips(ctx->seed, ctx->env, ctx->episode, g->num_players, flips);
    ctx->q = 0;
    g->shuffle = sk_rng_reshuffle;
    g->shuffle_ctx = ctx;
    return sk_reset_injected(g, deck, flips);
}

int64_t sk_rollout(int num_players, double score_penalty, int indirect, uint64_t seed,
                   uint64_t env0, int64_t games, uint64_t *checksum, double *score_sum) {
    sk_game g;
    if (sk_init(&g, num_players, score_penalty, indirect) != SK_OK) return -1;
    const int D = sk_obs_len(num_players, indirect);
    int8_t obs[19 + 12 * SK_MAX_PLAYERS], mask[26];
    int64_t steps = 0;
    uint64_t cs = 0;
    double ss = 0.0;
    for (int64_t e = 0; e < games; ++e) {
        sk_rng_shuffle_ctx ctx = {seed, env0 + (uint64_t)e, 0, 0};
        if (sk_reset_rng(&g, &ctx) != SK_OK) return -1;
        uint64_t t = 0;
        while (!g.is_terminated) { /* sample_game.py:10-21 */
            int pid = g.exp_player;
            sk_collect_observation(&g, pid, obs, mask);
            for (int i = 0; i < D; ++i) cs = cs * 1099511628211ull + (uint8_t)obs[i];
            for (int i = 0; i < 26; ++i) cs = cs * 1099511628211ull + (uint8_t)mask[i];
            int a = sk_rng_policy(seed, ctx.env, t++, mask);
            if (sk_act(&g, pid, a) < 0) return -2;
            ++steps;
        }
        for (int p = 0; p < num_players; ++p) ss += g.final_score[p];
    }
    if (checksum) *checksum = cs;
    if (score_sum) *score_sum = ss;
    return steps;
}
