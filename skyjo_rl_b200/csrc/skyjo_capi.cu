// skyjo_capi.cu -- the C ABI of libskyjo_b200.so (include/skyjo_b200.h): handle management,
// launch scheduling of the deal / step / observe kernels, statistics, host RNG twins.
// No torch, no CPU fallback: every compute entry launches CUDA kernels or fails.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <new>
#include <vector>

#include "../../include/skyjo_b200.h"
#include "skyjo_deal.cuh"
#include "skyjo_hostio.cuh"
#include "skyjo_hostsimd.h"
#include "skyjo_policy.h"
#include "skyjo_rng.cuh"
#include "skyjo_sample.cuh"
#include "skyjo_state.cuh"
#include "skyjo_step.cuh"

namespace skyjo {
// development builds (build.build_variant) instantiate only some player counts: the others stay
// undefined weak symbols (null) and skyjo_create rejects them
#ifdef SKYJO_ONLY_PLAYERS_MASK
#define SKYJO_WEAK __attribute__((weak))
#else
#define SKYJO_WEAK
#endif
#define SKYJO_DECL(n)                                                                                     \
    SKYJO_WEAK cudaError_t launch_step_##n(const StepParams &, bool, bool, cudaStream_t);                 \
    SKYJO_WEAK cudaError_t launch_rollout_##n(const StepParams &, const RolloutParams &, bool, cudaStream_t); \
    SKYJO_WEAK cudaError_t launch_observe_##n(const StepParams &, bool, int, int8_t *, int8_t *, int, int, cudaStream_t);
SKYJO_DECL(1) SKYJO_DECL(2) SKYJO_DECL(3) SKYJO_DECL(4) SKYJO_DECL(5) SKYJO_DECL(6)
SKYJO_DECL(7) SKYJO_DECL(8) SKYJO_DECL(9) SKYJO_DECL(10) SKYJO_DECL(11) SKYJO_DECL(12)
#undef SKYJO_DECL

static const step_launch_fn kStep[SKYJO_MAX_PLAYERS] = {
    launch_step_1, launch_step_2, launch_step_3, launch_step_4,  launch_step_5,  launch_step_6,
    launch_step_7, launch_step_8, launch_step_9, launch_step_10, launch_step_11, launch_step_12};
static const rollout_launch_fn kRollout[SKYJO_MAX_PLAYERS] = {
    launch_rollout_1, launch_rollout_2, launch_rollout_3, launch_rollout_4,  launch_rollout_5,  launch_rollout_6,
    launch_rollout_7, launch_rollout_8, launch_rollout_9, launch_rollout_10, launch_rollout_11, launch_rollout_12};
static const observe_launch_fn kObserve[SKYJO_MAX_PLAYERS] = {
    launch_observe_1, launch_observe_2, launch_observe_3, launch_observe_4,  launch_observe_5,  launch_observe_6,
    launch_observe_7, launch_observe_8, launch_observe_9, launch_observe_10, launch_observe_11, launch_observe_12};
}  // namespace skyjo

namespace skyjo {
void expand_obs_records_portable(const uint8_t *rec, long long e0, long long e1, int D, int8_t *obs) {
    expand_obs_records(rec, e0, e1, D, obs);
}
}  // namespace skyjo

using namespace skyjo;

#ifndef SKYJO_DEFAULT_PF_DIST
#define SKYJO_DEFAULT_PF_DIST 1184
#endif

struct SkyjoHandle {
    SkyjoConfig cfg;
    int device;
    long long B, Bpad;
    unsigned long long seed, first_env;
    DeviceState st;
    long long *stats_tmp;  // device int64[NUM_STATS]
    // skyjo_stats_allreduce_async: local sums (double-buffered) reduced on the caller's stream, the collective on
    // stats_stream; ev_ar_done[k] = the all-reduce that read stats_ar[k] has finished
    long long *stats_ar;   // device int64[2][NUM_STATS]
    bool ar_ready, ar_pending[2];
    int ar_slot;
    cudaStream_t stats_stream;
    cudaEvent_t ev_ar_ready, ev_ar_done[2];
    SkyjoOutputs outs;
    bool bound;
    int bulk_ok;
    unsigned long long t;   // lockstep counter
    long long launches;
    int steps_since_deal;
    // Refill windows (see open_window / close_window): the envs that finish during window w are
    // flagged in needs_deal array w & 1 and re-dealt by one flagged deal launch when the window
    // closes; with the in-kernel policy that launch runs on deal_stream, concurrently with window w + 1.
    uint8_t *flags_base;  // [2][Bpad]
    int parity;
    bool deal_async_ready, deal_pending[2], deal_async_enabled;
    cudaStream_t deal_stream;
    cudaEvent_t ev_window, ev_deal[2];
    // env ranges stepped on their own streams by skyjo_step_random (see step_random_ranges)
    int n_ranges;
    bool ranges_ready;
    cudaStream_t range_stream[HOSTIO_MAX_CHUNKS];
    cudaEvent_t ev_fork, ev_join[HOSTIO_MAX_CHUNKS];
    // skyjo_step_random(n) replayed from a CUDA graph (see step_random_ranges): one instantiated graph per
    // (n, starting flag-array parity); t_dev = the device-resident lockstep counter its kernels read
    struct StepGraph {
        int n_steps, parity, seen;
        long long launches;
        cudaGraphExec_t exec;
    };
    std::vector<StepGraph> graphs;
    bool graphs_enabled;
    unsigned long long *t_dev, t_dev_val;
    bool t_dev_valid;
    long long graph_replays;
    cudaStream_t capture_stream;
    int obs_len;
    int pf_dist;  // L2 prefetch distance of the step kernel, in tiles
    // optional per-kernel event timing (skyjo_step_random_profile)
    bool profiling;
    // Profiling mode: CUDA-event pairs on the launching stream.  A pair brackets one deal / rollout
    // launch, or one WINDOW of back-to-back single-step launches (up to 8, the launches between two
    // refill deals) -- no event between them, because an event record between two 45 us kernels
    // adds ~6 us to each pair (measured against ncu's gpu__time_duration) and breaks their
    // programmatic dependent launch, i.e. it times something the step loop never runs.
    std::vector<cudaEvent_t> prof_events;  // pairs (begin, end)
    std::vector<int> prof_kind;            // 0 step kernel, 1 deal kernel
    std::vector<int> prof_count;           // launches the pair covers
    bool prof_window_open = false;
    // skyjo_step_host wire staging (skyjo_hostio.cuh), created on first use
    bool hostio_ready;
    cudaStream_t copy_stream;
    cudaEvent_t ev_chunk_ready[HOSTIO_MAX_CHUNKS], ev_small_done[HOSTIO_MAX_CHUNKS], ev_counter_done, ev_all_done;
    uint32_t *packed_dev, *packed_host;     // [B]
    unsigned int *counter_dev, *counter_host;   // [2]: finished envs of the call, "row not encodable" flag
    uint8_t *rec_dev, *rec_host;            // compact observation records, [B][obs_record_bytes(D)]
    cudaEvent_t ev_rec_done[HOSTIO_MAX_CHUNKS];
    uint8_t *small_host, *small_dev;        // host-mapped staging of the small-batch path (B <= SMALL_HOST_MAX)
    unsigned int small_seq;
    int wire_mode;                          // 0 raw rows, 1 compact records, 2 mixed (adaptive share of the env ranges)
    // mixed mode: mix_k of 8 env ranges travel as compact records.  The share is tuned by measurement: the mean
    // call time of every window of 6 calls is filed under its k; the next window runs at the untried neighbour of
    // the fastest k seen, or stays there for mix_hold windows once both neighbours are known to be slower.
    int mix_k, mix_calls, mix_hold;
    double mix_acc_us, mix_t[9];
    long long last_d2h_bytes;               // bytes queued device -> host by the last skyjo_step_host
    double *entries_host, *entries_dev;     // host-mapped pinned, [cap][1 + N]
    unsigned int sparse_cap;
    int host_threads;
    HostPool *pool;
    double trace_us[6];                     // SKYJO_HOSTIO_TRACE: enqueue, wait small, expand, scatter, wait all
    long long trace_calls;
    const double *last_reward_host;         // buffer whose non-zero rows are tracked
    std::vector<long long> last_reward_rows;
};

static void prof_begin(SkyjoHandle *h, int kind, cudaStream_t s) {
    if (!h->profiling) return;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    h->prof_events.push_back(a);
    h->prof_events.push_back(b);
    h->prof_kind.push_back(kind);
    h->prof_count.push_back(1);
    cudaEventRecord(a, s);
}
static void prof_end(SkyjoHandle *h, cudaStream_t s) {
    if (!h->profiling) return;
    cudaEventRecord(h->prof_events.back(), s);
}
static void prof_close_window(SkyjoHandle *h, cudaStream_t s) {
    if (!h->profiling || !h->prof_window_open) return;
    cudaEventRecord(h->prof_events.back(), s);
    h->prof_window_open = false;
}

static thread_local char g_err[512] = "";

static int fail(int code, const char *fmt, const char *detail = "") {
    snprintf(g_err, sizeof(g_err), fmt, detail);
    return code;
}
static int cuda_fail(cudaError_t e, const char *where) {
    snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
    return SKYJO_E_CUDA + (int)e;
}
#define CU(call)                                      \
    do {                                              \
        cudaError_t _e = (call);                      \
        if (_e != cudaSuccess) return cuda_fail(_e, #call); \
    } while (0)

static long long align_up(long long x, long long a) { return (x + a - 1) / a * a; }

static bool config_ok(const SkyjoConfig *c) {
    return c && c->num_players >= 1 && c->num_players <= SKYJO_MAX_PLAYERS &&  // skyjo.py:24-26
           c->max_episode_steps >= 0 && c->max_episode_steps <= 0xFFFF && c->auto_reset >= SKYJO_RESET_OFF &&
           c->auto_reset <= SKYJO_RESET_NEXT_STEP;
}

struct Layout {
    long long planes, next_planes, pile, episode, needs_deal, stats, stats_tmp, stats_ar, errflag, total;
};
static Layout layout_for(int N, long long B) {
    const long long Bpad = align_up(B, ENV_PAD);
    const long long np = num_planes(N);
    Layout L;
    long long off = 0;
    L.planes = off;      off = align_up(off + np * Bpad * 16, 256);
    L.next_planes = off; off = align_up(off + np * Bpad * 16, 256);
    L.pile = off;        off = align_up(off + 2 * Bpad * PILE_ROW, 256);
    L.episode = off;     off = align_up(off + Bpad * 4, 256);
    L.needs_deal = off;  off = align_up(off + 2 * Bpad, 256);
    L.stats = off;       off = align_up(off + (long long)STAT_SLOTS * NUM_STATS * 8, 256);
    L.stats_tmp = off;   off = align_up(off + NUM_STATS * 8, 256);
    L.stats_ar = off;    off = align_up(off + 2 * NUM_STATS * 8, 256);
    L.errflag = off;     off = align_up(off + 4, 256);
    L.total = off;
    return L;
}

extern "C" {

int skyjo_abi_version(void) { return SKYJO_ABI_VERSION; }
const char *skyjo_last_error(void) { return g_err; }

int skyjo_obs_len(const SkyjoConfig *cfg) {
    if (!config_ok(cfg)) return -1;
    return cfg->observe_other_player_indirect ? 31 : 19 + 12 * cfg->num_players;  // skyjo.py:43-45
}

int64_t skyjo_state_bytes(const SkyjoConfig *cfg, int64_t num_envs) {
    if (!config_ok(cfg) || num_envs <= 0) return -1;
    return layout_for(cfg->num_players, num_envs).total;
}

int skyjo_create(const SkyjoConfig *cfg, int device, int64_t num_envs, uint64_t seed, int64_t first_global_env_id,
                 void *state_dev, int64_t state_bytes, SkyjoHandle **out) {
    if (!out) return fail(SKYJO_E_INVALID, "null out pointer");
    *out = nullptr;
    if (!config_ok(cfg)) return fail(SKYJO_E_INVALID, "invalid config: num_players must be 1..12, auto_reset one of SKYJO_RESET_*");
    if (num_envs <= 0 || first_global_env_id < 0) return fail(SKYJO_E_INVALID, "num_envs must be > 0");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(SKYJO_E_NO_DEVICE, "no CUDA device: libskyjo_b200 has no CPU fallback");
    }
    if (device < 0 || device >= ndev) return fail(SKYJO_E_INVALID, "bad device index");
    if (!kStep[cfg->num_players - 1] || !kObserve[cfg->num_players - 1] || !kRollout[cfg->num_players - 1])
        return fail(SKYJO_E_INVALID, "this development build does not instantiate that player count");
    const Layout L = layout_for(cfg->num_players, num_envs);
    if (!state_dev || state_bytes < L.total || ((uintptr_t)state_dev & 255)) {
        return fail(SKYJO_E_INVALID, "state buffer null, too small or not 256-byte aligned");
    }
    CU(cudaSetDevice(device));
    if (const char *g = getenv("SKYJO_L2_FETCH")) {  // experiment knob: L2 fetch granularity hint (32/64/128)
        cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(g));
        cudaGetLastError();
    }
    SkyjoHandle *h = new (std::nothrow) SkyjoHandle();
    if (!h) return fail(SKYJO_E_INVALID, "out of host memory");
    h->cfg = *cfg;
    h->device = device;
    h->B = num_envs;
    h->Bpad = align_up(num_envs, ENV_PAD);
    h->seed = seed;
    h->first_env = (unsigned long long)first_global_env_id;
    uint8_t *base = (uint8_t *)state_dev;
    h->st.planes = (U128 *)(base + L.planes);
    h->st.next_planes = (U128 *)(base + L.next_planes);
    h->st.deck = base + L.pile;
    h->st.episode = (uint32_t *)(base + L.episode);
    h->st.needs_deal = base + L.needs_deal;
    h->flags_base = base + L.needs_deal;
    h->parity = 0;
    h->deal_async_ready = false;
    h->deal_pending[0] = h->deal_pending[1] = false;
    h->deal_async_enabled = getenv("SKYJO_SYNC_DEAL") == nullptr;
    h->ranges_ready = false;
    h->graphs_enabled = getenv("SKYJO_NO_GRAPH") == nullptr;
    h->t_dev = nullptr;
    h->graph_replays = 0;
    h->capture_stream = nullptr;
    h->t_dev_val = 0;
    h->t_dev_valid = false;
    h->n_ranges = num_envs >= (1 << 18) ? 4 : 1;
    if (const char *g = getenv("SKYJO_RANGES")) h->n_ranges = atoi(g);  // experiment knob
    if (h->n_ranges < 1) h->n_ranges = 1;
    if (h->n_ranges > HOSTIO_MAX_CHUNKS) h->n_ranges = HOSTIO_MAX_CHUNKS;
    h->st.stats = (unsigned long long *)(base + L.stats);
    h->stats_tmp = (long long *)(base + L.stats_tmp);
    h->stats_ar = (long long *)(base + L.stats_ar);
    h->ar_ready = false;
    h->ar_pending[0] = h->ar_pending[1] = false;
    h->ar_slot = 0;
    h->st.errflag = (uint32_t *)(base + L.errflag);
    h->bound = false;
    h->bulk_ok = 0;
    h->t = 0;
    h->launches = 0;
    h->steps_since_deal = 0;
    h->obs_len = skyjo_obs_len(cfg);
    h->profiling = false;
    h->hostio_ready = false;
    h->pool = nullptr;
    h->host_threads = 0;
    // default: mixed where the CPU has the streaming expansions (csrc/skyjo_hostsimd.cpp), else raw rows
    h->wire_mode = host_simd_level() >= 2 ? 2 : 0;
    if (const char *g = getenv("SKYJO_HOST_WIRE"))
        h->wire_mode = (atoi(g) == 1 || !strcmp(g, "compact")) ? 1 : ((atoi(g) == 2 || !strcmp(g, "mixed")) ? 2 : 0);
    {
        // starting point of the measured share: what it converged to on the B200 boxes of round 2 -- 4 of 8 ranges with
        // the host to itself, none when 4 or more ranks share the host's memory system
        int ranks = 1;
        if (const char *g = getenv("LOCAL_WORLD_SIZE")) ranks = atoi(g) > 0 ? atoi(g) : 1;
        h->mix_k = ranks == 1 ? 4 : (ranks < 4 ? 2 : 0);
    }
    if (const char *g = getenv("SKYJO_HOST_MIX")) h->mix_k = atoi(g);   // experiment knob: fixes the share
    h->mix_calls = -1;  // the first call allocates: not timed
    h->mix_hold = 0;
    h->mix_acc_us = 0.0;
    for (double &v : h->mix_t) v = -1.0;
    h->last_d2h_bytes = 0;
    h->rec_dev = h->rec_host = nullptr;
    h->small_host = h->small_dev = nullptr;
    h->small_seq = 0;
    h->last_reward_host = nullptr;
    h->trace_calls = 0;
    for (double &v : h->trace_us) v = 0.0;
    h->pf_dist = SKYJO_DEFAULT_PF_DIST;
    if (const char *g = getenv("SKYJO_PF_DIST")) h->pf_dist = atoi(g);  // experiment knob
    cudaError_t e = cudaMemset(state_dev, 0, (size_t)L.total);
    if (e != cudaSuccess) {
        delete h;
        return cuda_fail(e, "cudaMemset(state)");
    }
    *out = h;
    return SKYJO_OK;
}

static void hostio_release(SkyjoHandle *h) {
    if (h->trace_calls && getenv("SKYJO_HOSTIO_TRACE")) {
        const double n = (double)h->trace_calls;
        fprintf(stderr, "[skyjo_step_host] %lld calls, mean us: enqueue %.1f | wait packed %.1f | expand %.1f | "
                        "scatter rewards %.1f | wait obs %.1f\n", h->trace_calls, h->trace_us[0] / n, h->trace_us[1] / n,
                h->trace_us[2] / n, h->trace_us[3] / n, h->trace_us[4] / n);
    }
    delete h->pool;
    h->pool = nullptr;
    if (h->small_host) {
        cudaSetDevice(h->device);
        cudaDeviceSynchronize();
        cudaFreeHost(h->small_host);
        h->small_host = h->small_dev = nullptr;
    }
    if (!h->hostio_ready) return;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->copy_stream);
    cudaFree(h->packed_dev);
    cudaFree(h->counter_dev);
    cudaFreeHost(h->packed_host);
    cudaFreeHost(h->counter_host);
    cudaFreeHost(h->entries_host);
    cudaFree(h->rec_dev);
    cudaFreeHost(h->rec_host);
    h->rec_dev = h->rec_host = nullptr;
    for (int c = 0; c < HOSTIO_MAX_CHUNKS; ++c) {
        cudaEventDestroy(h->ev_chunk_ready[c]);
        cudaEventDestroy(h->ev_small_done[c]);
        cudaEventDestroy(h->ev_rec_done[c]);
    }
    cudaEventDestroy(h->ev_counter_done);
    cudaEventDestroy(h->ev_all_done);
    cudaStreamDestroy(h->copy_stream);
    h->hostio_ready = false;
}

// the instantiated step graphs bake the output pointers, the seed and the env ranges: dropped when one changes
static void graphs_drop(SkyjoHandle *h) {
    for (auto &g : h->graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    h->graphs.clear();
}

int skyjo_destroy(SkyjoHandle *h) {
    if (h && (!h->graphs.empty() || h->t_dev)) {
        cudaSetDevice(h->device);
        cudaDeviceSynchronize();
        graphs_drop(h);
        cudaFree(h->t_dev);
        if (h->capture_stream) cudaStreamDestroy(h->capture_stream);
    }
    if (h && h->ranges_ready) {
        cudaSetDevice(h->device);
        for (int r = 0; r < HOSTIO_MAX_CHUNKS; ++r) {
            cudaStreamSynchronize(h->range_stream[r]);
            cudaStreamDestroy(h->range_stream[r]);
            cudaEventDestroy(h->ev_join[r]);
        }
        cudaEventDestroy(h->ev_fork);
    }
    if (h && h->deal_async_ready) {
        cudaSetDevice(h->device);
        cudaStreamSynchronize(h->deal_stream);
        cudaEventDestroy(h->ev_window);
        cudaEventDestroy(h->ev_deal[0]);
        cudaEventDestroy(h->ev_deal[1]);
        cudaStreamDestroy(h->deal_stream);
    }
    if (h && h->ar_ready) {
        cudaSetDevice(h->device);
        cudaStreamSynchronize(h->stats_stream);
        cudaEventDestroy(h->ev_ar_ready);
        cudaEventDestroy(h->ev_ar_done[0]);
        cudaEventDestroy(h->ev_ar_done[1]);
        cudaStreamDestroy(h->stats_stream);
    }
    if (h) hostio_release(h);
    delete h;
    return SKYJO_OK;
}

int skyjo_bind_outputs(SkyjoHandle *h, const SkyjoOutputs *o) {
    if (!h || !o) return fail(SKYJO_E_INVALID, "null argument");
    if (!o->obs_dev || !o->action_mask_dev || !o->agent_dev || !o->done_dev || !o->reward_dev || !o->final_score_dev)
        return fail(SKYJO_E_INVALID, "all six output buffers are required");
    if (((uintptr_t)o->reward_dev & 7) || ((uintptr_t)o->final_score_dev & 7))
        return fail(SKYJO_E_INVALID, "reward / final_score must be 8-byte aligned");
    graphs_drop(h);
    h->outs = *o;
    h->bulk_ok = (((uintptr_t)o->obs_dev & 15) == 0 && ((uintptr_t)o->action_mask_dev & 15) == 0) ? 1 : 0;
    h->bound = true;
    return SKYJO_OK;
}

static StepParams make_params(const SkyjoHandle *h) {
    StepParams p;
    memset(&p, 0, sizeof(p));
    p.st = h->st;
    p.obs = (int8_t *)h->outs.obs_dev;
    p.mask = (int8_t *)h->outs.action_mask_dev;
    p.agent = (int8_t *)h->outs.agent_dev;
    p.done = (uint8_t *)h->outs.done_dev;
    p.reward = (double *)h->outs.reward_dev;
    p.final_score = (double *)h->outs.final_score_dev;
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
    p.bulk_ok = h->bulk_ok;
    p.pdl = getenv("SKYJO_NO_PDL") ? 0 : 1;
    p.pf_dist = h->pf_dist;
    return p;
}

static int launch_deal(SkyjoHandle *h, int flagged, int target_next, const int8_t *decks, const uint8_t *flips,
                       cudaStream_t s, long long e_begin = 0, long long e_end = -1) {
    DealParams d;
    d.st = h->st;
    d.B = e_end < 0 ? h->B : e_end;
    d.e_begin = flagged ? e_begin : 0;
    d.Bpad = h->Bpad;
    d.first_env = h->first_env;
    d.seed = h->seed;
    d.N = h->cfg.num_players;
    d.indirect = h->cfg.observe_other_player_indirect ? 1 : 0;
    d.flagged = flagged;
    d.target_next = target_next;
    d.decks = decks;
    d.flips = flips;
    const long long per = flagged ? DEAL_SCAN : DEAL_THREADS;
    const unsigned grid = (unsigned)((d.B - d.e_begin + per - 1) / per);
    prof_begin(h, 1, s);
    deal_kernel<<<grid, DEAL_THREADS, 0, s>>>(d);
    prof_end(h, s);
    h->launches += 1;
    CU(cudaGetLastError());
    return SKYJO_OK;
}

// ---- refill windows -----------------------------------------------------------------------------
// A window is a run of consecutive steps (at most deal_period of them) whose finished envs are
// flagged in needs_deal array `parity`.  Closing it launches the flagged deal D_w that pre-deals
// the next-but-one episode of those envs and clears their flags.  An episode lasts >= 21 act()
// calls under legal play, so an env that finished in window w cannot finish again in window w + 1:
// D_w touches nothing window w + 1 reads (the env's next_planes row, its free deck slot, its flag
// in the other array) and may run concurrently with it on deal_stream.  It has to be complete
// before window w + 2 starts, which also is the next user of flag array w & 1.
static int deal_async_init(SkyjoHandle *h) {
    if (h->deal_async_ready) return SKYJO_OK;
    CU(cudaStreamCreateWithFlags(&h->deal_stream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&h->ev_window, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&h->ev_deal[0], cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&h->ev_deal[1], cudaEventDisableTiming));
    h->deal_async_ready = true;
    return SKYJO_OK;
}

// before the first kernel of a window: the deal that last used this window's flag array is done
static int open_window(SkyjoHandle *h, cudaStream_t s) {
    if (h->steps_since_deal == 0 && h->deal_pending[h->parity]) {
        CU(cudaStreamWaitEvent(s, h->ev_deal[h->parity], 0));
        h->deal_pending[h->parity] = false;
    }
    return SKYJO_OK;
}

static int close_window(SkyjoHandle *h, cudaStream_t s, bool async) {
    h->steps_since_deal = 0;
    if (!h->cfg.auto_reset) return SKYJO_OK;
    int rc;
    if (async && h->deal_async_enabled && !h->profiling) {  // per-kernel event timing measures each kernel alone
        rc = deal_async_init(h);
        if (rc) return rc;
        CU(cudaEventRecord(h->ev_window, s));
        CU(cudaStreamWaitEvent(h->deal_stream, h->ev_window, 0));
        rc = launch_deal(h, 1, 1, nullptr, nullptr, h->deal_stream);
        if (rc) return rc;
        CU(cudaEventRecord(h->ev_deal[h->parity], h->deal_stream));
        h->deal_pending[h->parity] = true;
    } else {
        rc = launch_deal(h, 1, 1, nullptr, nullptr, s);
        if (rc) return rc;
    }
    h->parity ^= 1;
    h->st.needs_deal = h->flags_base + (size_t)h->parity * (size_t)h->Bpad;
    return SKYJO_OK;
}

// `s` waits for every deal still running on deal_stream (before anything else touches the state)
static int join_deals(SkyjoHandle *h, cudaStream_t s) {
    for (int k = 0; k < 2; ++k)
        if (h->deal_pending[k]) {
            CU(cudaStreamWaitEvent(s, h->ev_deal[k], 0));
            h->deal_pending[k] = false;
        }
    return SKYJO_OK;
}

// close the running window (if any steps are in it) and join: afterwards no flag is set
static int quiesce(SkyjoHandle *h, cudaStream_t s) {
    if (h->steps_since_deal > 0) {
        int rc = close_window(h, s, false);
        if (rc) return rc;
    }
    return join_deals(h, s);
}

static int reset_common(SkyjoHandle *h, const int8_t *decks, const uint8_t *flips, cudaStream_t s) {
    if (!h) return fail(SKYJO_E_INVALID, "null handle");
    if (!h->bound) return fail(SKYJO_E_NOT_BOUND, "call skyjo_bind_outputs first");
    CU(cudaSetDevice(h->device));
    int rc = join_deals(h, s);
    if (rc) return rc;
    h->parity = 0;
    h->st.needs_deal = h->flags_base;
    CU(cudaMemsetAsync(h->flags_base, 0, 2 * (size_t)h->Bpad, s));
    rc = launch_deal(h, 0, 0, decks, flips, s);
    if (rc) return rc;
    if (h->cfg.auto_reset) {
        rc = launch_deal(h, 0, 1, nullptr, nullptr, s);
        if (rc) return rc;
    }
    h->steps_since_deal = 0;
    StepParams p = make_params(h);
    CU(kObserve[h->cfg.num_players - 1](p, h->cfg.observe_other_player_indirect != 0, -1, p.obs, p.mask, 1,
                                        h->bulk_ok, s));
    h->launches += 1;
    return SKYJO_OK;
}

int skyjo_reset(SkyjoHandle *h, void *stream) { return reset_common(h, nullptr, nullptr, (cudaStream_t)stream); }

int skyjo_reset_injected(SkyjoHandle *h, const int8_t *decks_dev, const uint8_t *flips_dev, void *stream) {
    if (!decks_dev || !flips_dev) return fail(SKYJO_E_INVALID, "decks and flips are required");
    return reset_common(h, decks_dev, flips_dev, (cudaStream_t)stream);
}

int skyjo_seed(SkyjoHandle *h, uint64_t seed, void *stream) {
    if (!h) return fail(SKYJO_E_INVALID, "null handle");
    CU(cudaSetDevice(h->device));
    // a refill deal still running on deal_stream reads and bumps episode[]: it has to be over before the
    // counters are zeroed, or the flagged envs restart at episode 1 / 2 instead of 0
    int rc = quiesce(h, (cudaStream_t)stream);
    if (rc) return rc;
    graphs_drop(h);
    h->seed = seed;
    h->t = 0;
    CU(cudaMemsetAsync(h->st.episode, 0, (size_t)h->Bpad * 4, (cudaStream_t)stream));
    return skyjo_reset(h, stream);
}

// Refill cadence of the pre-dealt "next" episodes: the window length P in steps.  The deal D_w of the envs that
// finished in window w has to be complete before any of them can finish AGAIN, and it is only awaited when window
// w + 2 opens, so 2 P must not exceed the shortest possible episode.  Under legal play (the in-kernel policy) a
// seat uncovers at most one of its ten hidden cards per turn (skyjo.py:389-404) and the game ends at the draw
// action of a seat with none left (:350-356): the starter needs 9 N + 1 turns of two act() calls, the other N - 1
// seats then play once more, and the terminating draw call follows -- 20 N + 1 act() calls at least (21 for one
// player, 81 for four).  P = the largest multiple of 8 below half of that, at most 32: 8 / 16 / 24 / 32 steps for
// N = 1 / 2 / 3 / >= 4.  Longer windows mean fewer, fuller flagged deal launches (a deal CTA compacts the hits of
// 1024 envs: 61 decks per CTA at P = 8 and N = 4, i.e. two warps; 240 at P = 32).  A step cap shortens episodes;
// external actions may be illegal and end an episode at once, so they deal after every step.
static int deal_period(const SkyjoHandle *h, bool policy) {
    if (!policy) return 1;
    int shortest = 20 * h->cfg.num_players + 1;
    if (h->cfg.max_episode_steps > 0 && h->cfg.max_episode_steps < shortest) shortest = h->cfg.max_episode_steps;
    int p = shortest / 2 / 8 * 8;
    if (const char *g = getenv("SKYJO_DEAL_PERIOD")) p = atoi(g) < p ? atoi(g) : p;  // experiment knob (may only shorten)
    return p < 8 ? 1 : (p > 32 ? 32 : p);
}

static int step_once(SkyjoHandle *h, const void *actions, int dtype, bool policy, cudaStream_t s) {
    int rc = open_window(h, s);
    if (rc) return rc;
    StepParams p = make_params(h);
    p.actions = actions;
    p.action_dtype = dtype;
    if (h->profiling) {
        if (!h->prof_window_open) {
            prof_begin(h, 0, s);
            h->prof_window_open = true;
        } else {
            h->prof_count.back() += 1;
        }
    }
    cudaError_t le = kStep[h->cfg.num_players - 1](p, h->cfg.observe_other_player_indirect != 0, policy, s);
    CU(le);
    h->launches += 1;
    h->t += 1;
    if (++h->steps_since_deal >= deal_period(h, policy)) {
        prof_close_window(h, s);
        return close_window(h, s, policy);
    }
    return SKYJO_OK;
}

int skyjo_step(SkyjoHandle *h, const void *actions_dev, int action_dtype, void *stream) {
    if (!h || !actions_dev) return fail(SKYJO_E_INVALID, "null argument");
    if (action_dtype < SKYJO_ACT_U8 || action_dtype > SKYJO_ACT_I64) return fail(SKYJO_E_INVALID, "bad action dtype");
    if (!h->bound) return fail(SKYJO_E_NOT_BOUND, "call skyjo_bind_outputs first");
    CU(cudaSetDevice(h->device));
    // external actions may be illegal and end an episode at once: every step is its own window,
    // refilled in stream order
    int rc = quiesce(h, (cudaStream_t)stream);
    if (rc) return rc;
    return step_once(h, actions_dev, action_dtype, false, (cudaStream_t)stream);
}

__global__ void set_t_kernel(unsigned long long *t_dev, unsigned long long v) { *t_dev = v; }
__global__ void advance_t_kernel(unsigned long long *t_dev, unsigned long long n) { *t_dev += n; }

// number of refill deals n lockstep steps with the in-kernel policy end with (the last, shorter window included)
static int windows_in(const SkyjoHandle *h, int n_steps) {
    const int period = deal_period(h, true);
    return (n_steps + period - 1) / period;
}

// The launches of step_random_ranges replay unchanged from call to call when (a) the lockstep counter comes from
// device memory and (b) the call toggles the flag-array parity an even number of times: then one instantiated CUDA
// graph per (n_steps, starting parity) stands for the ~4 n kernel launches and their fork / join events, and the
// host side of a call shrinks from ~1 ms per 64 steps (3.7 us per launch; as long as the device takes at 2^18
// envs, and what eight ranks sharing 32 cores stumble over) to one cudaGraphLaunch.
static bool graph_eligible(const SkyjoHandle *h, int n_steps) {
    return h->graphs_enabled && !h->profiling && n_steps >= 8 && (!h->cfg.auto_reset || windows_in(h, n_steps) % 2 == 0);
}

// enqueues the n_steps x ranges step launches and their flagged deals, forked from / joined back into `s`; with
// `graph` the kernels take their lockstep counter as *t_dev + offset (the calls are being captured)
static int enqueue_ranges(SkyjoHandle *h, int n_steps, cudaStream_t s, bool graph) {
    const int R = h->n_ranges;
    const long long per = align_up((h->B + R - 1) / R, ENV_PAD);
    long long b0[HOSTIO_MAX_CHUNKS], b1[HOSTIO_MAX_CHUNKS];
    int nr = 0;
    for (long long b = 0; b < h->B; b += per) {
        b0[nr] = b;
        b1[nr] = b + per < h->B ? b + per : h->B;
        ++nr;
    }
    CU(cudaEventRecord(h->ev_fork, s));
    for (int r = 0; r < nr; ++r) CU(cudaStreamWaitEvent(h->range_stream[r], h->ev_fork, 0));
    const int period = deal_period(h, true);
    const bool ind = h->cfg.observe_other_player_indirect != 0;
    int in_window = 0;
    for (int i = 0; i < n_steps; ++i) {
        StepParams p = make_params(h);
        if (graph) {
            p.t_base = h->t_dev;
            p.t = (unsigned long long)i;
        }
        for (int r = 0; r < nr; ++r) {
            p.tile_off = b0[r] / TILE;
            p.tiles = (b1[r] - b0[r] + TILE - 1) / TILE;
            CU(kStep[h->cfg.num_players - 1](p, ind, true, h->range_stream[r]));
            h->launches += 1;
        }
        h->t += 1;
        if (++in_window >= period || i + 1 == n_steps) {
            in_window = 0;
            if (h->cfg.auto_reset) {
                for (int r = 0; r < nr; ++r) {
                    int rc = launch_deal(h, 1, 1, nullptr, nullptr, h->range_stream[r], b0[r], b1[r]);
                    if (rc) return rc;
                }
                h->parity ^= 1;
                h->st.needs_deal = h->flags_base + (size_t)h->parity * (size_t)h->Bpad;
            }
        }
    }
    for (int r = 0; r < nr; ++r) {
        CU(cudaEventRecord(h->ev_join[r], h->range_stream[r]));
        CU(cudaStreamWaitEvent(s, h->ev_join[r], 0));
    }
    return SKYJO_OK;
}

// The envs of a batch never interact, so a batch can be stepped as R independent env ranges, each
// on its own stream: step kernel (tile sub-range) x n, with the range's flagged deal after every
// window, all in stream order.  The GPU interleaves the ranges, so the launch ramp and tail of one
// range's kernel are covered by the other ranges' CTAs (at 2^20 envs a lone full-batch launch
// loses ~7 of 52 us to them), without any cross-kernel memory-ordering assumption.
static int step_random_ranges(SkyjoHandle *h, int n_steps, cudaStream_t s) {
    if (!h->ranges_ready) {
        for (int r = 0; r < HOSTIO_MAX_CHUNKS; ++r) {
            CU(cudaStreamCreateWithFlags(&h->range_stream[r], cudaStreamNonBlocking));
            CU(cudaEventCreateWithFlags(&h->ev_join[r], cudaEventDisableTiming));
        }
        CU(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
        h->ranges_ready = true;
    }
    int rc = quiesce(h, s);  // no window open, no deal pending on deal_stream
    if (rc) return rc;
    if (!graph_eligible(h, n_steps)) return enqueue_ranges(h, n_steps, s, false);

    // a call shape is captured the second time it is seen (capture + instantiation cost about two direct calls)
    SkyjoHandle::StepGraph *g = nullptr;
    for (auto &c : h->graphs)
        if (c.n_steps == n_steps && c.parity == h->parity) g = &c;
    if (!g) {
        if (h->graphs.size() >= 16) {  // a caller cycling through many shapes: forget the oldest
            if (h->graphs.front().exec) cudaGraphExecDestroy(h->graphs.front().exec);
            h->graphs.erase(h->graphs.begin());
        }
        h->graphs.push_back({n_steps, h->parity, 1, 0, nullptr});
        return enqueue_ranges(h, n_steps, s, false);
    }
    if (!g->exec) {
        if (!h->t_dev) CU(cudaMalloc(&h->t_dev, 8));
        const unsigned long long t0 = h->t;
        const long long l0 = h->launches;
        const int parity0 = h->parity;
        cudaGraph_t graph = nullptr;
        // captured on a stream of the library's own (the caller's may be the legacy default stream, which cannot
        // capture); the graph is launched into the caller's stream
        if (!h->capture_stream) CU(cudaStreamCreateWithFlags(&h->capture_stream, cudaStreamNonBlocking));
        const cudaStream_t cs = h->capture_stream;
        CU(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
        rc = enqueue_ranges(h, n_steps, cs, true);
        if (rc == SKYJO_OK) {
            advance_t_kernel<<<1, 1, 0, cs>>>(h->t_dev, (unsigned long long)n_steps);
            if (cudaGetLastError() != cudaSuccess) rc = SKYJO_E_CUDA;
        }
        const cudaError_t ce = cudaStreamEndCapture(cs, &graph);
        // the capture ran the host bookkeeping of one call without executing anything: take it back
        g->launches = h->launches - l0 + 1;
        h->t = t0;
        h->launches = l0;
        h->parity = parity0;
        h->st.needs_deal = h->flags_base + (size_t)h->parity * (size_t)h->Bpad;
        cudaError_t ie = cudaErrorUnknown;
        if (rc == SKYJO_OK && ce == cudaSuccess && graph) ie = cudaGraphInstantiate(&g->exec, graph, 0);
        if (graph) cudaGraphDestroy(graph);
        if (ie != cudaSuccess) {  // no graphs on this driver / in this context: the direct path from now on
            cudaGetLastError();
            g->exec = nullptr;
            h->graphs_enabled = false;
            return enqueue_ranges(h, n_steps, s, false);
        }
    }
    if (!h->t_dev_valid || h->t_dev_val != h->t) {
        set_t_kernel<<<1, 1, 0, s>>>(h->t_dev, h->t);
        CU(cudaGetLastError());
    }
    CU(cudaGraphLaunch(g->exec, s));
    h->t += (unsigned long long)n_steps;
    h->t_dev_val = h->t;
    h->t_dev_valid = true;
    h->launches += g->launches;
    h->graph_replays += 1;
    return SKYJO_OK;
}

int skyjo_step_random(SkyjoHandle *h, int n_steps, void *stream) {
    if (!h || n_steps < 0) return fail(SKYJO_E_INVALID, "bad argument");
    if (!h->bound) return fail(SKYJO_E_NOT_BOUND, "call skyjo_bind_outputs first");
    CU(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;
    // one range: the graph pays where the loop is launch-bound (a 2^15-env step kernel takes about the 3.7 us its
    // launch costs the host; at 2^16 the async refill deals of the loop below are worth more)
    if (!h->profiling && n_steps >= 2 && (h->n_ranges > 1 || (h->B <= (1 << 15) && graph_eligible(h, n_steps))))
        return step_random_ranges(h, n_steps, s);
    for (int i = 0; i < n_steps; ++i) {
        int rc = step_once(h, nullptr, 0, true, s);
        if (rc) return rc;
    }
    // every call ends with its window closed (the deal may still be running on deal_stream)
    prof_close_window(h, s);
    if (h->steps_since_deal > 0) return close_window(h, s, true);
    return SKYJO_OK;
}

int skyjo_set_env_ranges(SkyjoHandle *h, int n) {
    if (!h || n < 0 || n > HOSTIO_MAX_CHUNKS) return fail(SKYJO_E_INVALID, "env ranges must be 0 (default) .. 8");
    graphs_drop(h);
    h->n_ranges = n > 0 ? n : (h->B >= (1 << 18) ? 4 : 1);
    return SKYJO_OK;
}

// Multi-step launches with the in-kernel policy (rollout_kernel, skyjo_step.cuh): n_steps env-steps
// in ceil(n_steps / K) launches, K = the refill cadence, each followed by one flagged deal launch.
int skyjo_rollout_random(SkyjoHandle *h, int n_steps, const SkyjoRollout *out, void *stream) {
    if (!h || n_steps < 0 || !out) return fail(SKYJO_E_INVALID, "bad argument");
    if (!out->obs_dev || !out->action_mask_dev || !out->agent_dev || !out->done_dev)
        return fail(SKYJO_E_INVALID, "all four rollout buffers are required");
    if (!h->bound) return fail(SKYJO_E_NOT_BOUND, "call skyjo_bind_outputs first");
    CU(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;
    const long long B = h->B, D = h->obs_len;
    const int K = h->cfg.auto_reset ? deal_period(h, true) : 8;
    if (h->steps_since_deal > 0) {  // an open window of single steps: close it first
        int rc = close_window(h, s, true);
        if (rc) return rc;
    }
    const bool aligned = (((uintptr_t)out->obs_dev | (uintptr_t)out->action_mask_dev) & 15) == 0 &&
                         (B * D) % 16 == 0 && (B * 26) % 16 == 0;
    for (int t0 = 0; t0 < n_steps; t0 += K) {
        int rc = open_window(h, s);
        if (rc) return rc;
        StepParams p = make_params(h);
        RolloutParams r;
        r.K = n_steps - t0 < K ? n_steps - t0 : K;
        r.obs = (int8_t *)out->obs_dev + (long long)t0 * B * D;
        r.mask = (int8_t *)out->action_mask_dev + (long long)t0 * B * 26;
        r.agent = (int8_t *)out->agent_dev + (long long)t0 * B;
        r.done = (uint8_t *)out->done_dev + (long long)t0 * B;
        r.publish = t0 + r.K == n_steps ? 1 : 0;
        r.bulk_ok = aligned ? 1 : 0;
        prof_begin(h, 0, s);
        cudaError_t le = kRollout[h->cfg.num_players - 1](p, r, h->cfg.observe_other_player_indirect != 0, s);
        prof_end(h, s);
        CU(le);
        h->launches += 1;
        h->t += (unsigned long long)r.K;
        rc = close_window(h, s, true);
        if (rc) return rc;
    }
    return SKYJO_OK;
}

int skyjo_profile_begin(SkyjoHandle *h) {
    if (!h) return fail(SKYJO_E_INVALID, "null handle");
    h->profiling = true;
    return SKYJO_OK;
}

int skyjo_profile_end(SkyjoHandle *h, void *stream, double *step_ms, double *deal_ms, int64_t *n_step_launches,
                      int64_t *n_deal_launches) {
    if (!h || !step_ms || !deal_ms) return fail(SKYJO_E_INVALID, "null argument");
    h->profiling = false;
    cudaError_t se = cudaStreamSynchronize((cudaStream_t)stream);
    double sum[2] = {0.0, 0.0};
    long long cnt[2] = {0, 0};
    for (size_t i = 0; i < h->prof_kind.size(); ++i) {
        float ms = 0.f;
        if (se == cudaSuccess && cudaEventElapsedTime(&ms, h->prof_events[2 * i], h->prof_events[2 * i + 1]) == cudaSuccess) {
            sum[h->prof_kind[i]] += ms;
            cnt[h->prof_kind[i]] += h->prof_count[i];
        }
        cudaEventDestroy(h->prof_events[2 * i]);
        cudaEventDestroy(h->prof_events[2 * i + 1]);
    }
    h->prof_events.clear();
    h->prof_kind.clear();
    h->prof_count.clear();
    h->prof_window_open = false;
    CU(se);
    *step_ms = sum[0];
    *deal_ms = sum[1];
    if (n_step_launches) *n_step_launches = cnt[0];
    if (n_deal_launches) *n_deal_launches = cnt[1];
    return SKYJO_OK;
}

int skyjo_step_random_profile(SkyjoHandle *h, int n_steps, void *stream, double *step_ms, double *deal_ms,
                              int64_t *n_step_launches, int64_t *n_deal_launches) {
    if (!h || !step_ms || !deal_ms) return fail(SKYJO_E_INVALID, "null argument");
    h->profiling = true;
    const int rc = skyjo_step_random(h, n_steps, stream);
    const int rc2 = skyjo_profile_end(h, stream, step_ms, deal_ms, n_step_launches, n_deal_launches);
    return rc ? rc : rc2;
}

static int default_host_threads() {
    // this rank's share of the CPUs the process may run on (ranks of one box split them evenly), at most 16:
    // the expansion is memory-bound long before that
    int n = (int)HostPool::rank_cpus().size();
    if (n < 1) {
        int hw = (int)std::thread::hardware_concurrency();
        int ranks = 1;  // ranks sharing this host under torchrun
        if (const char *g = getenv("LOCAL_WORLD_SIZE")) ranks = atoi(g) > 0 ? atoi(g) : 1;
        n = hw / ranks;
    }
    if (const char *g = getenv("SKYJO_HOST_THREADS")) return atoi(g) < 1 ? 1 : (atoi(g) > 64 ? 64 : atoi(g));
    n -= 1;  // one core of the slice stays free for the caller's other threads (CUDA's, the interpreter's)
    return n < 2 ? 2 : (n > 8 ? 8 : n);
}

int skyjo_set_host_threads(SkyjoHandle *h, int n) {
    if (!h || n < 0 || n > 64) return fail(SKYJO_E_INVALID, "host threads must be 0 (default) .. 64");
    h->host_threads = n;
    delete h->pool;
    h->pool = nullptr;
    return SKYJO_OK;
}

static int hostio_init(SkyjoHandle *h) {
    if (h->hostio_ready) return SKYJO_OK;
    const size_t B = (size_t)h->B, N = (size_t)h->cfg.num_players;
    h->sparse_cap = (unsigned int)(B / 8 > 64 ? B / 8 : 64);
    CU(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    for (int c = 0; c < HOSTIO_MAX_CHUNKS; ++c) {
        CU(cudaEventCreateWithFlags(&h->ev_chunk_ready[c], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&h->ev_small_done[c], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&h->ev_rec_done[c], cudaEventDisableTiming));
    }
    CU(cudaEventCreateWithFlags(&h->ev_counter_done, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&h->ev_all_done, cudaEventDisableTiming));
    CU(cudaMalloc(&h->packed_dev, B * 4));
    CU(cudaMalloc(&h->counter_dev, 8));
    CU(cudaHostAlloc(&h->packed_host, B * 4, cudaHostAllocDefault));
    CU(cudaHostAlloc(&h->counter_host, 8, cudaHostAllocDefault));
    CU(cudaHostAlloc(&h->entries_host, (size_t)h->sparse_cap * (1 + N) * 8, cudaHostAllocMapped));
    CU(cudaHostGetDevicePointer(&h->entries_dev, h->entries_host, 0));
    h->hostio_ready = true;
    return SKYJO_OK;
}

// skyjo_step_host for B <= SMALL_HOST_MAX envs: latency path (skyjo_hostio.cuh, publish_small_kernel)
static int step_host_small(SkyjoHandle *h, const uint8_t *actions_host, int8_t *obs_host, int8_t *mask_host,
                           int8_t *agent_host, uint8_t *done_host, double *reward_host, cudaStream_t s) {
    const size_t B = (size_t)h->B, N = (size_t)h->cfg.num_players, D = (size_t)h->obs_len;
    const SmallLayout L = small_layout(B, D, N);
    if (!h->small_host) {
        CU(cudaHostAlloc(&h->small_host, L.total, cudaHostAllocMapped));
        CU(cudaHostGetDevicePointer(&h->small_dev, h->small_host, 0));
        memset(h->small_host, 0, L.total);
    }
    int rc = quiesce(h, s);
    if (rc) return rc;
    memcpy(h->small_host + L.act, actions_host, B);
    StepParams p = make_params(h);
    p.actions = h->small_dev + L.act;
    p.action_dtype = SKYJO_ACT_U8;
    CU(kStep[h->cfg.num_players - 1](p, h->cfg.observe_other_player_indirect != 0, false, s));
    const unsigned int seq = ++h->small_seq ? h->small_seq : ++h->small_seq;  // never 0
    publish_small_kernel<<<1, 256, 0, s>>>((const int8_t *)h->outs.obs_dev, (const int8_t *)h->outs.action_mask_dev,
                                           (const int8_t *)h->outs.agent_dev, (const uint8_t *)h->outs.done_dev,
                                           (const double *)h->outs.reward_dev, h->B, (int)D, (int)N, h->small_dev, seq);
    CU(cudaGetLastError());
    h->launches += 2;
    h->t += 1;
    // poll the sequence word; past ~2 ms something is wrong: let the stream report it
    volatile unsigned int *flag = reinterpret_cast<volatile unsigned int *>(h->small_host);
    const auto t0 = std::chrono::steady_clock::now();
    while (flag[0] != seq) {
#if defined(__x86_64__)
        _mm_pause();
#endif
        if (std::chrono::steady_clock::now() - t0 > std::chrono::milliseconds(2)) {
            CU(cudaStreamSynchronize(s));
            if (flag[0] != seq) return fail(SKYJO_E_STATE, "skyjo_step_host: the outputs never arrived in host memory");
        }
    }
    if (obs_host) memcpy(obs_host, h->small_host + L.obs, B * D);
    if (mask_host) memcpy(mask_host, h->small_host + L.mask, B * 26);
    if (agent_host) memcpy(agent_host, h->small_host + L.agent, B);
    if (done_host) memcpy(done_host, h->small_host + L.done, B);
    if (reward_host) memcpy(reward_host, h->small_host + L.reward, B * N * 8);
    h->last_reward_host = nullptr;
    h->last_d2h_bytes = (long long)(L.total - L.obs);
    // the refill deal only when an episode ended (external actions may end one at any step)
    if (flag[1] != 0u) {
        h->steps_since_deal = 1;
        return close_window(h, s, false);
    }
    h->steps_since_deal = 0;
    return SKYJO_OK;
}

int skyjo_step_host(SkyjoHandle *h, const uint8_t *actions_host, int8_t *obs_host, int8_t *mask_host,
                    int8_t *agent_host, uint8_t *done_host, double *reward_host, void *stream) {
    if (!h || !actions_host) return fail(SKYJO_E_INVALID, "null argument");
    if (!h->bound) return fail(SKYJO_E_NOT_BOUND, "call skyjo_bind_outputs first");
    CU(cudaSetDevice(h->device));
    if (h->B <= SMALL_HOST_MAX && !getenv("SKYJO_HOST_NO_SMALL"))
        return step_host_small(h, actions_host, obs_host, mask_host, agent_host, done_host, reward_host, (cudaStream_t)stream);
    int rc = hostio_init(h);
    if (rc) return rc;
    if (!h->pool) h->pool = new (std::nothrow) HostPool(h->host_threads > 0 ? h->host_threads : default_host_threads());
    if (!h->pool) return fail(SKYJO_E_INVALID, "out of host memory");
    cudaStream_t s = (cudaStream_t)stream, cs = h->copy_stream;
    const size_t B = (size_t)h->B, N = (size_t)h->cfg.num_players;
    using clk = std::chrono::steady_clock;
    auto t_prev = clk::now();
    const auto t_call0 = t_prev;
    auto lap = [&](int k) {
        const auto now = clk::now();
        h->trace_us[k] += std::chrono::duration<double, std::micro>(now - t_prev).count();
        t_prev = now;
    };
    // the agent buffer doubles as the staging area of the uint8 actions: it is rewritten by the step
    uint8_t *act_dev = (uint8_t *)h->outs.agent_dev;
    rc = quiesce(h, s);
    if (rc) return rc;
    const bool want_small = mask_host || agent_host || done_host || reward_host;
    const int RB = obs_record_bytes(h->obs_len);
    // env ranges: multiples of ENV_PAD envs, so every range starts on a tile and on a 16-byte boundary
    int chunks = h->B >= (1 << 18) ? ((obs_host && h->wire_mode != 0) ? 8 : 4) : (h->B >= (1 << 16) ? 2 : 1);
    if (const char *g = getenv("SKYJO_HOST_CHUNKS")) chunks = atoi(g);
    if (chunks < 1) chunks = 1;
    if (chunks > HOSTIO_MAX_CHUNKS) chunks = HOSTIO_MAX_CHUNKS;
    const long long per = align_up((h->B + chunks - 1) / chunks, ENV_PAD);
    long long c_begin[HOSTIO_MAX_CHUNKS], c_end[HOSTIO_MAX_CHUNKS];
    int nc = 0;
    for (long long b = 0; b < h->B; b += per) {
        c_begin[nc] = b;
        c_end[nc] = b + per < h->B ? b + per : h->B;
        ++nc;
    }
    // Per range, the observation rows travel raw (the copy engine writes them into obs_host, no CPU work) or as
    // compact records (1.7x fewer bytes on the link, expanded by the workers).  Mixed mode sends the first
    // n_compact ranges compact -- their expansion then overlaps the raw copies of the later ranges -- and moves
    // the share towards the faster call time (perturb and observe: what limits the call -- the link, or the host's
    // memory system that the copy engine and the workers share -- differs from box to box).
    const bool adapt = h->wire_mode == 2 && !getenv("SKYJO_HOST_MIX");
    int n_compact = !obs_host ? 0 : (h->wire_mode == 1 ? nc : (h->wire_mode == 2 ? (h->mix_k * nc + 4) / 8 : 0));
    if (n_compact > nc) n_compact = nc;
    if (n_compact < 0) n_compact = 0;
    const bool any_compact = n_compact > 0;
    if (obs_host && h->wire_mode != 0 && !h->rec_dev) {  // record staging, in the first call that may need it later
        CU(cudaMalloc(&h->rec_dev, B * (size_t)RB));
        CU(cudaHostAlloc(&h->rec_host, B * (size_t)RB + 256, cudaHostAllocDefault));  // + the gather windows' over-read
    }
    if (want_small || any_compact) CU(cudaMemsetAsync(h->counter_dev, 0, 8, s));
    StepParams p = make_params(h);
    p.actions = act_dev;
    p.action_dtype = SKYJO_ACT_U8;
    long long d2h = 0;
    for (int c = 0; c < nc; ++c) {
        const bool compact = c < n_compact;
        const size_t e0 = (size_t)c_begin[c], n = (size_t)(c_end[c] - c_begin[c]);
        CU(cudaMemcpyAsync(act_dev + e0, actions_host + e0, n, cudaMemcpyHostToDevice, s));
        p.tile_off = c_begin[c] / TILE;
        p.tiles = (long long)((n + TILE - 1) / TILE);
        CU(kStep[h->cfg.num_players - 1](p, h->cfg.observe_other_player_indirect != 0, false, s));
        h->launches += 1;
        if (want_small) {
            pack_host_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(
                h->st.planes, h->Bpad, c_begin[c], c_end[c], (int)N, (const uint8_t *)h->outs.done_dev,
                (const double *)h->outs.reward_dev, h->packed_dev, h->counter_dev,
                reward_host ? h->entries_dev : nullptr, h->sparse_cap);
            h->launches += 1;
            CU(cudaGetLastError());
        }
        if (compact) {
            compact_obs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(
                (const int8_t *)h->outs.obs_dev, c_begin[c], c_end[c], h->obs_len, RB, h->rec_dev, h->counter_dev + 1);
            h->launches += 1;
            CU(cudaGetLastError());
        }
        CU(cudaEventRecord(h->ev_chunk_ready[c], s));
        CU(cudaStreamWaitEvent(cs, h->ev_chunk_ready[c], 0));
        if (want_small) {
            CU(cudaMemcpyAsync(h->packed_host + e0, h->packed_dev + e0, n * 4, cudaMemcpyDeviceToHost, cs));
            CU(cudaEventRecord(h->ev_small_done[c], cs));
            d2h += (long long)n * 4;
        }
        if (compact) {
            CU(cudaMemcpyAsync(h->rec_host + e0 * (size_t)RB, h->rec_dev + e0 * (size_t)RB, n * (size_t)RB,
                               cudaMemcpyDeviceToHost, cs));
            CU(cudaEventRecord(h->ev_rec_done[c], cs));
            d2h += (long long)n * RB;
        } else if (obs_host) {
            CU(cudaMemcpyAsync(obs_host + e0 * (size_t)h->obs_len, (const int8_t *)h->outs.obs_dev + e0 * (size_t)h->obs_len,
                               n * (size_t)h->obs_len, cudaMemcpyDeviceToHost, cs));
            d2h += (long long)n * h->obs_len;
        }
    }
    h->t += 1;
    if (reward_host || any_compact) {
        CU(cudaMemcpyAsync(h->counter_host, h->counter_dev, 8, cudaMemcpyDeviceToHost, cs));
        CU(cudaEventRecord(h->ev_counter_done, cs));
        d2h += 8;
    }
    CU(cudaEventRecord(h->ev_all_done, cs));
    // external actions may end any episode at once: refill after every step (queued behind the last
    // pack kernel, it overlaps the copies)
    h->steps_since_deal = 1;
    rc = close_window(h, s, false);
    if (rc) return rc;
    lap(0);

    // previous call's reward rows back to zero while the first chunk travels
    if (reward_host) {
        if (h->last_reward_host != reward_host) {
            memset(reward_host, 0, B * N * 8);
            h->last_reward_host = reward_host;
        } else {
            for (long long e : h->last_reward_rows) memset(reward_host + (size_t)e * N, 0, N * 8);
        }
        h->last_reward_rows.clear();
    }
    if (want_small || any_compact) {
        const uint32_t *packed = h->packed_host;
        const uint8_t *rec = h->rec_host;
        const int D = h->obs_len;
        for (int c = 0; c < nc; ++c) {
            const bool compact = c < n_compact;
            if (!want_small && !compact) continue;
            // the record copy of a chunk is queued behind its packed words: one wait covers both
            CU(cudaEventSynchronize(compact ? h->ev_rec_done[c] : h->ev_small_done[c]));
            lap(1);
            const long long b0 = c_begin[c], nB = c_end[c] - c_begin[c];
            h->pool->run([=](int part, int parts) {
                // parts are multiples of 64 envs (b0 is one): whole cache lines of every output per worker
                const long long g = (nB + 63) / 64;
                const long long e0 = b0 + g * part / parts * 64;
                const long long e1 = part + 1 == parts ? b0 + nB : b0 + g * (part + 1) / parts * 64;
                if (e0 >= e1) return;
                if (want_small) {
                    if (mask_host && agent_host && done_host)
                        expand_packed_wide(packed, e0, e1, mask_host, agent_host, done_host);
                    else
                        expand_packed(packed, e0, e1, mask_host, agent_host, done_host);
                }
                if (compact) expand_obs_records_wide(rec, e0, e1, D, obs_host, 256);
            });
            lap(2);
        }
        if (reward_host || any_compact) CU(cudaEventSynchronize(h->ev_counter_done));
        if (reward_host) {
            const unsigned int cnt = h->counter_host[0];
            if (cnt <= h->sparse_cap) {
                // every pack kernel finished before ev_counter_done: its writes to the mapped buffer are visible
                for (unsigned int i = 0; i < cnt; ++i) {
                    const double *src = h->entries_host + (size_t)i * (1 + N);
                    const long long e = (long long)reinterpret_cast<const unsigned long long *>(src)[0];
                    memcpy(reward_host + (size_t)e * N, src + 1, N * 8);
                    h->last_reward_rows.push_back(e);
                }
            } else {
                // more finished envs than the compact buffer holds: dense copy of the reward tensor
                CU(cudaMemcpyAsync(reward_host, h->outs.reward_dev, B * N * 8, cudaMemcpyDeviceToHost, cs));
                CU(cudaEventRecord(h->ev_all_done, cs));
                d2h += (long long)(B * N * 8);
                h->last_reward_host = nullptr;  // every row may be non-zero: start from a memset next time
            }
        }
        if (any_compact && h->counter_host[1] != 0u) {
            // a row outside the record's symbol sets (not reachable from the 4-bit-bin state layout): dense copy
            CU(cudaMemcpyAsync(obs_host, h->outs.obs_dev, B * (size_t)D, cudaMemcpyDeviceToHost, cs));
            CU(cudaEventRecord(h->ev_all_done, cs));
            d2h += (long long)(B * (size_t)D);
        }
    }
    lap(3);
    CU(cudaEventSynchronize(h->ev_all_done));
    lap(4);
    h->trace_calls += 1;
    h->last_d2h_bytes = d2h;
    if (adapt && obs_host && nc >= 4) {
        if (h->mix_calls >= 0) h->mix_acc_us += std::chrono::duration<double, std::micro>(clk::now() - t_call0).count();
        if (++h->mix_calls >= 6) {
            const double t = h->mix_acc_us / h->mix_calls;
            int k = h->mix_k < 0 ? 0 : (h->mix_k > 8 ? 8 : h->mix_k);
            h->mix_t[k] = h->mix_t[k] < 0.0 ? t : 0.5 * (h->mix_t[k] + t);
            if (h->mix_hold > 0) {
                if (--h->mix_hold == 0)  // look around again: the neighbours' figures are stale
                    for (int q = 0; q <= 8; ++q)
                        if (q != k) h->mix_t[q] = -1.0;
            } else {
                int best = k;
                for (int q = 0; q <= 8; ++q)
                    if (h->mix_t[q] >= 0.0 && h->mix_t[q] < h->mix_t[best]) best = q;
                if (best > 0 && h->mix_t[best - 1] < 0.0) k = best - 1;
                else if (best < 8 && h->mix_t[best + 1] < 0.0) k = best + 1;
                else {
                    k = best;
                    h->mix_hold = 40;
                }
                h->mix_k = k;
            }
            h->mix_calls = 0;
            h->mix_acc_us = 0.0;
        }
    }
    return SKYJO_OK;
}

int skyjo_set_host_wire(SkyjoHandle *h, int mode) {
    if (!h || mode < 0 || mode > 2)
        return fail(SKYJO_E_INVALID, "wire mode must be 0 (raw rows), 1 (compact records) or 2 (mixed, adaptive)");
    h->wire_mode = mode;
    return SKYJO_OK;
}

int64_t skyjo_host_wire_bytes(const SkyjoHandle *h) { return h ? h->last_d2h_bytes : -1; }
int skyjo_host_wire_share(const SkyjoHandle *h) { return h ? (h->wire_mode == 1 ? 8 : (h->wire_mode == 2 ? h->mix_k : 0)) : -1; }

int skyjo_observe(SkyjoHandle *h, int agent, void *obs_dev, void *mask_dev, void *stream) {
    if (!h || !obs_dev || !mask_dev) return fail(SKYJO_E_INVALID, "null argument");
    if (agent >= h->cfg.num_players) return fail(SKYJO_E_INVALID, "agent out of range");
    CU(cudaSetDevice(h->device));
    StepParams p = make_params(h);
    const int bulk = (((uintptr_t)obs_dev & 15) == 0 && ((uintptr_t)mask_dev & 15) == 0) ? 1 : 0;
    CU(kObserve[h->cfg.num_players - 1](p, h->cfg.observe_other_player_indirect != 0, agent, (int8_t *)obs_dev,
                                        (int8_t *)mask_dev, 0, bulk, (cudaStream_t)stream));
    h->launches += 1;
    return SKYJO_OK;
}

int skyjo_stats_device(SkyjoHandle *h, int64_t *out_dev, void *stream) {
    if (!h || !out_dev) return fail(SKYJO_E_INVALID, "null argument");
    CU(cudaSetDevice(h->device));
    stats_reduce_kernel<<<1, NUM_STATS, 0, (cudaStream_t)stream>>>(h->st.stats, (long long *)out_dev);
    h->launches += 1;
    CU(cudaGetLastError());
    return SKYJO_OK;
}

// The one collective of the library: sum of the statistics vector over the ranks of an NCCL
// communicator, off the step path.  NCCL is not a link-time dependency: the symbols are taken from whichever
// libnccl the process has loaded (torch's bundled one under torch.distributed, or one the caller dlopen'ed with
// RTLD_GLOBAL), so the communicator and the call always belong to the same NCCL build.
typedef int (*nccl_allreduce_fn)(const void *, void *, size_t, int, int, void *, cudaStream_t);
typedef const char *(*nccl_errstr_fn)(int);
static nccl_allreduce_fn g_nccl_all_reduce = nullptr;
static nccl_errstr_fn g_nccl_err_string = nullptr;

static int nccl_resolve() {
    if (!g_nccl_all_reduce) {
        g_nccl_all_reduce = (nccl_allreduce_fn)dlsym(RTLD_DEFAULT, "ncclAllReduce");
        g_nccl_err_string = (nccl_errstr_fn)dlsym(RTLD_DEFAULT, "ncclGetErrorString");
        if (!g_nccl_all_reduce) {
            if (void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL | RTLD_NOLOAD)) {  // loaded RTLD_LOCAL
                g_nccl_all_reduce = (nccl_allreduce_fn)dlsym(lib, "ncclAllReduce");
                g_nccl_err_string = (nccl_errstr_fn)dlsym(lib, "ncclGetErrorString");
            }
        }
    }
    if (!g_nccl_all_reduce) return fail(SKYJO_E_INVALID, "NCCL is not loaded in this process (ncclAllReduce not found)");
    return SKYJO_OK;
}

static int nccl_sum_i64(const void *src, void *dst, void *comm, cudaStream_t s) {
    const int ncclInt64 = 4, ncclSum = 0;  // nccl.h: ncclDataType_t / ncclRedOp_t
    const int nrc = g_nccl_all_reduce(src, dst, (size_t)NUM_STATS, ncclInt64, ncclSum, comm, s);
    if (nrc != 0)
        return fail(SKYJO_E_NCCL, "ncclAllReduce failed: %s", g_nccl_err_string ? g_nccl_err_string(nrc) : "unknown NCCL error");
    return SKYJO_OK;
}

int skyjo_stats_allreduce(SkyjoHandle *h, void *nccl_comm, int64_t *out_dev, void *stream) {
    if (!h || !nccl_comm || !out_dev) return fail(SKYJO_E_INVALID, "null argument");
    int rc = nccl_resolve();
    if (rc) return rc;
    rc = skyjo_stats_device(h, out_dev, stream);
    if (rc) return rc;
    return nccl_sum_i64(out_dev, out_dev, nccl_comm, (cudaStream_t)stream);
}

// The same collective OFF the caller's stream ("on a side stream after a device-side reduction").  On
// `stream` only the one-CTA reduction of the 256 replicated counter vectors into stats_ar[k] runs (a snapshot at
// this point of the stream, ~3 us); the ncclAllReduce -- a rendezvous with every other rank -- runs on the
// library's stats stream behind an event, so the step launches queued on `stream` afterwards never wait for a
// peer.  The two stats_ar buffers alternate: buffer k is rewritten two calls later, after `stream` has waited for
// the collective that read it (long finished by then).
int skyjo_stats_allreduce_async(SkyjoHandle *h, void *nccl_comm, int64_t *out_dev, void *stream) {
    if (!h || !out_dev) return fail(SKYJO_E_INVALID, "null argument");
    CU(cudaSetDevice(h->device));
    int rc = nccl_comm ? nccl_resolve() : SKYJO_OK;
    if (rc) return rc;
    if (!h->ar_ready) {
        CU(cudaStreamCreateWithFlags(&h->stats_stream, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&h->ev_ar_ready, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&h->ev_ar_done[0], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&h->ev_ar_done[1], cudaEventDisableTiming));
        h->ar_ready = true;
    }
    cudaStream_t s = (cudaStream_t)stream, ss = h->stats_stream;
    const int k = h->ar_slot;
    if (h->ar_pending[k]) {
        CU(cudaStreamWaitEvent(s, h->ev_ar_done[k], 0));
        h->ar_pending[k] = false;
    }
    long long *local = h->stats_ar + (size_t)k * NUM_STATS;
    rc = skyjo_stats_device(h, (int64_t *)local, stream);
    if (rc) return rc;
    CU(cudaEventRecord(h->ev_ar_ready, s));
    CU(cudaStreamWaitEvent(ss, h->ev_ar_ready, 0));
    if (nccl_comm) {
        rc = nccl_sum_i64(local, out_dev, nccl_comm, ss);
        if (rc) return rc;
    } else {  // one rank: the sum is the local vector
        CU(cudaMemcpyAsync(out_dev, local, NUM_STATS * 8, cudaMemcpyDeviceToDevice, ss));
    }
    CU(cudaEventRecord(h->ev_ar_done[k], ss));
    h->ar_pending[k] = true;
    h->ar_slot = k ^ 1;
    return SKYJO_OK;
}

int skyjo_stats_allreduce_wait(SkyjoHandle *h, void *stream) {
    if (!h) return fail(SKYJO_E_INVALID, "null handle");
    CU(cudaSetDevice(h->device));
    for (int k = 0; k < 2; ++k)
        if (h->ar_pending[k]) {
            CU(cudaStreamWaitEvent((cudaStream_t)stream, h->ev_ar_done[k], 0));
            h->ar_pending[k] = false;
        }
    return SKYJO_OK;
}

int skyjo_stats_host(SkyjoHandle *h, int64_t *out_host, void *stream) {
    if (!h || !out_host) return fail(SKYJO_E_INVALID, "null argument");
    int rc = skyjo_stats_device(h, (int64_t *)h->stats_tmp, stream);
    if (rc) return rc;
    CU(cudaMemcpyAsync(out_host, h->stats_tmp, NUM_STATS * 8, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CU(cudaStreamSynchronize((cudaStream_t)stream));
    return SKYJO_OK;
}

int skyjo_stats_clear(SkyjoHandle *h, void *stream) {
    if (!h) return fail(SKYJO_E_INVALID, "null handle");
    CU(cudaSetDevice(h->device));
    CU(cudaMemsetAsync(h->st.stats, 0, (size_t)STAT_SLOTS * NUM_STATS * 8, (cudaStream_t)stream));
    return SKYJO_OK;
}

int skyjo_sample_actions(SkyjoHandle *h, const float *logits_dev, const int8_t *mask_dev, uint64_t sample_seed,
                         uint8_t *actions_dev, float *logp_dev, float *entropy_dev, void *stream) {
    if (!h || !logits_dev || !actions_dev || !logp_dev) return fail(SKYJO_E_INVALID, "null argument");
    if (!mask_dev && !h->bound) return fail(SKYJO_E_NOT_BOUND, "no mask given and no outputs bound");
    if (((uintptr_t)logits_dev & 7) || ((uintptr_t)(mask_dev ? mask_dev : (const int8_t *)h->outs.action_mask_dev) & 1))
        return fail(SKYJO_E_INVALID, "logits must be 8-byte aligned, the mask 2-byte aligned");
    CU(cudaSetDevice(h->device));
    const int8_t *mask = mask_dev ? mask_dev : (const int8_t *)h->outs.action_mask_dev;
    sample_actions_kernel<<<(unsigned)((h->B + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        logits_dev, mask, h->B, h->first_env, sample_seed, h->t, actions_dev, logp_dev, entropy_dev);
    h->launches += 1;
    CU(cudaGetLastError());
    return SKYJO_OK;
}

// ---- fused policy forward (csrc/skyjo_policy.cu) ----------------------------------------------------------------
int64_t skyjo_policy_packed_bytes(void) { return POLICY_PACKED_BYTES; }

int skyjo_policy_pack(int obs_len, int n_out, const float *w1, const float *b1, const float *w2, const float *b2,
                      const float *w3, const float *b3, void *packed_dev, void *stream) {
    if (!w1 || !b1 || !w2 || !b2 || !w3 || !b3 || !packed_dev) return fail(SKYJO_E_INVALID, "null argument");
    if (obs_len < 1 || obs_len > POLICY_MAX_OBS)
        return fail(SKYJO_E_INVALID, "the fused policy kernel takes observation rows of at most 96 bytes (direct mode up to 6 players, or indirect)");
    if (n_out < 1 || n_out > 26) return fail(SKYJO_E_INVALID, "n_out must be 26 (logits) or 1 (value head)");
    if (((uintptr_t)packed_dev & 15) != 0) return fail(SKYJO_E_INVALID, "packed buffer must be 16-byte aligned");
    CU(launch_policy_pack(w1, b1, w2, b2, w3, b3, obs_len, n_out, packed_dev, (cudaStream_t)stream));
    return SKYJO_OK;
}

static int policy_launch(SkyjoHandle *h, const void *packed_dev, PolicyParams &p, void *stream) {
    if (!h || !packed_dev) return fail(SKYJO_E_INVALID, "null argument");
    if (!h->bound) return fail(SKYJO_E_NOT_BOUND, "call skyjo_bind_outputs first");
    if (h->obs_len > POLICY_MAX_OBS) return fail(SKYJO_E_INVALID, "observation rows longer than 96 bytes: use a library policy");
    if (((uintptr_t)packed_dev & 15) != 0) return fail(SKYJO_E_INVALID, "packed buffer must be 16-byte aligned");
    CU(cudaSetDevice(h->device));
    int sms = 0;
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device));
    p.obs = (const int8_t *)h->outs.obs_dev;
    p.mask = (const int8_t *)h->outs.action_mask_dev;
    p.packed = (const uint8_t *)packed_dev;
    p.B = h->B;
    p.D = h->obs_len;
    p.bulk_ok = h->bulk_ok;
    p.first_env = h->first_env;
    p.t = h->t;
    CU(launch_policy(p, sms, (cudaStream_t)stream));
    h->launches += 1;
    return SKYJO_OK;
}

int skyjo_policy_sample(SkyjoHandle *h, const void *packed_dev, uint64_t sample_seed, uint8_t *actions_dev,
                        float *logp_dev, float *entropy_dev, float *logits_dev, void *stream) {
    if (!actions_dev) return fail(SKYJO_E_INVALID, "null argument");
    PolicyParams p;
    memset(&p, 0, sizeof(p));
    p.seed = sample_seed;
    p.actions = actions_dev;
    p.logp = logp_dev;
    p.entropy = entropy_dev;
    p.logits = logits_dev;
    return policy_launch(h, packed_dev, p, stream);
}

int skyjo_policy_value(SkyjoHandle *h, const void *packed_dev, float *value_dev, void *stream) {
    if (!value_dev) return fail(SKYJO_E_INVALID, "null argument");
    PolicyParams p;
    memset(&p, 0, sizeof(p));
    p.value = value_dev;
    return policy_launch(h, packed_dev, p, stream);
}

int skyjo_policy_debug(SkyjoHandle *h, const void *packed_dev, float *pre1_dev, float *pre2_dev, float *logits_dev,
                       void *stream) {
    if (!logits_dev) return fail(SKYJO_E_INVALID, "null argument");
    PolicyParams p;
    memset(&p, 0, sizeof(p));
    p.logits = logits_dev;
    p.dbg1 = pre1_dev;
    p.dbg2 = pre2_dev;
    return policy_launch(h, packed_dev, p, stream);
}

int skyjo_policy_trace(SkyjoHandle *h, const void *packed_dev, uint8_t *actions_dev, int64_t *trace_dev, void *stream) {
    if (!actions_dev || !trace_dev) return fail(SKYJO_E_INVALID, "null argument");
    PolicyParams p;
    memset(&p, 0, sizeof(p));
    p.actions = actions_dev;
    p.trace = (long long *)trace_dev;
    return policy_launch(h, packed_dev, p, stream);
}

int skyjo_policy_trace_len(void) { return POLICY_TRACE_LEN; }

int skyjo_quiesce(SkyjoHandle *h, void *stream) {
    if (!h) return fail(SKYJO_E_INVALID, "null handle");
    CU(cudaSetDevice(h->device));
    return quiesce(h, (cudaStream_t)stream);
}

int skyjo_export_debug(SkyjoHandle *h, int64_t env0, int64_t count, SkyjoEnvDebug *out_dev, void *stream) {
    if (!h || !out_dev) return fail(SKYJO_E_INVALID, "null argument");
    if (env0 < 0 || count <= 0 || env0 + count > h->B) return fail(SKYJO_E_INVALID, "env range out of bounds");
    CU(cudaSetDevice(h->device));
    {
        int rc = join_deals(h, (cudaStream_t)stream);
        if (rc) return rc;
    }
    const unsigned grid = (unsigned)((count + 127) / 128);
    export_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(h->st, h->Bpad, h->cfg.num_players,
                                                          h->cfg.observe_other_player_indirect ? 1 : 0, env0, count,
                                                          out_dev);
    h->launches += 1;
    CU(cudaGetLastError());
    return SKYJO_OK;
}

int skyjo_check(SkyjoHandle *h, void *stream) {
    if (!h) return fail(SKYJO_E_INVALID, "null handle");
    CU(cudaSetDevice(h->device));
    {
        int rc = join_deals(h, (cudaStream_t)stream);
        if (rc) return rc;
    }
    uint32_t flag = 0;
    CU(cudaMemcpyAsync(&flag, h->st.errflag, 4, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CU(cudaStreamSynchronize((cudaStream_t)stream));
    if (flag) {
        snprintf(g_err, sizeof(g_err), "device consistency flag 0x%x:%s%s%s%s", flag,
                 (flag & ERR_NEXT_NOT_READY) ? " next episode was not dealt in time" : "",
                 (flag & ERR_BAD_DECK) ? " injected deck outside -2..12 or >15 copies of a value" : "",
                 (flag & ERR_BAD_FLIPS) ? " injected flips invalid" : "",
                 (flag & ERR_ASSIST) ? " warp assist missed a rare event" : "");
        return SKYJO_E_STATE;
    }
    return SKYJO_OK;
}

int64_t skyjo_step_count(const SkyjoHandle *h) { return h ? (int64_t)h->t : -1; }
int skyjo_set_step_count(SkyjoHandle *h, int64_t t) {
    if (!h || t < 0) return fail(SKYJO_E_INVALID, "bad argument");
    h->t = (unsigned long long)t;
    // the restored buffer was saved quiesced (no flag set): start a fresh window on flag array 0
    h->steps_since_deal = 0;
    h->parity = 0;
    h->st.needs_deal = h->flags_base;
    return SKYJO_OK;
}
int64_t skyjo_launch_count(const SkyjoHandle *h) { return h ? h->launches : -1; }
int64_t skyjo_graph_replay_count(const SkyjoHandle *h) { return h ? h->graph_replays : -1; }

// ---- host twins of the device RNG (same header, host compilation path) ---------------------
void skyjo_host_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    U4 c = {ctr[0], ctr[1], ctr[2], ctr[3]};
    U4 r = philox4x32_10(c, key[0], key[1]);
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

void skyjo_host_deck(uint64_t seed, uint64_t genv, uint32_t episode, int8_t out[SKYJO_DECK]) {
    for (int i = 0; i < SKYJO_DECK; ++i) out[i] = (int8_t)(i / 10 - 2);
    U4 blk = {0, 0, 0, 0};
    for (int i = SKYJO_DECK - 1; i >= 1; --i) {
        const int k = SKYJO_DECK - 1 - i;
        if ((k & 3) == 0) blk = rng_block(seed, genv, PURPOSE_DEAL, episode, (uint32_t)(k >> 2));
        const uint32_t r = (k & 3) == 0 ? blk.x : (k & 3) == 1 ? blk.y : (k & 3) == 2 ? blk.z : blk.w;
        const int j = (int)bounded(r, (uint32_t)(i + 1));
        const int8_t t = out[i];
        out[i] = out[j];
        out[j] = t;
    }
}

void skyjo_host_flips(uint64_t seed, uint64_t genv, uint32_t episode, int num_players, uint8_t *out) {
    for (int q = 0; q < num_players; ++q) {
        U4 r = rng_block(seed, genv, PURPOSE_FLIPS, episode, (uint32_t)q);
        uint32_t a = bounded(r.x, 12u), b = bounded(r.y, 11u);
        if (b >= a) b += 1u;
        out[2 * q] = (uint8_t)a;
        out[2 * q + 1] = (uint8_t)b;
    }
}

// The device never materialises the reshuffled pile: env_step samples the remaining multiset one card per
// draw (skyjo_core.cuh: rng_block(PURPOSE_RESHUFFLE, episode, q << 16 | remaining) -> hist_take).  This twin
// plays the same draws forward over the whole pile and lays them out as the python list the reference
// expects (last element = new discard card, skyjo.py:135-138).
int skyjo_host_reshuffle(uint64_t seed, uint64_t genv, uint32_t episode, uint32_t reshuffle_index, int8_t *pile,
                         int len) {
    if (!pile || len < 1 || len > 512) return fail(SKYJO_E_INVALID, "pile must hold 1..512 cards");
    int bins[15] = {0};
    for (int i = 0; i < len; ++i) {
        if (pile[i] < -2 || pile[i] > 12) return fail(SKYJO_E_INVALID, "pile holds a card outside -2..12");
        bins[pile[i] + 2] += 1;
    }
    const uint32_t q = reshuffle_index & (uint32_t)HDR_Q_MASK;  // the header keeps 7 bits of the reshuffle index
    for (int d = 0; d < len; ++d) {
        const uint32_t remaining = (uint32_t)(len - d);
        const U4 r = rng_block(seed, genv, PURPOSE_RESHUFFLE, episode, (q << 16) | remaining);
        uint32_t idx = bounded(r.x, remaining);
        int code = 0;  // smallest code with bins[0] + .. + bins[code] > idx (hist_take)
        while (idx >= (uint32_t)bins[code]) idx -= (uint32_t)bins[code++];
        bins[code] -= 1;
        pile[len - 1 - d] = (int8_t)(code - 2);
    }
    return SKYJO_OK;
}

void skyjo_host_expand_packed(const uint32_t *packed, int64_t n, int8_t *mask, int8_t *agent, uint8_t *done) {
    if (mask && agent && done)
        expand_packed_wide(packed, 0, n, mask, agent, done);  // AVX-512 + streaming stores where the CPU has them
    else
        expand_packed(packed, 0, n, mask, agent, done);
}

int skyjo_host_obs_record_bytes(int obs_len) { return obs_len >= 31 && (obs_len - 19) % 12 == 0 ? obs_record_bytes(obs_len) : -1; }

int64_t skyjo_host_pack_obs(const int8_t *obs, int64_t n, int obs_len, uint8_t *rec) {
    const int RB = obs_record_bytes(obs_len);
    int64_t bad = 0;
    for (int64_t e = 0; e < n; ++e) bad += pack_obs_record(obs + e * obs_len, obs_len, rec + e * RB) ? 0 : 1;
    return bad;
}

void skyjo_host_expand_obs(const uint8_t *rec, int64_t n, int obs_len, int8_t *obs, int portable) {
    if (portable == 1)
        expand_obs_scalar(rec, 0, n, obs_len, obs);
    else if (portable == 2)
        expand_obs_records(rec, 0, n, obs_len, obs);       // 16-byte shuffles, plain stores
    else if (portable == 3)
        expand_obs_records_wide(rec, 0, n, obs_len, obs, 0);    // L1 staging + streaming stores where available
    else
        // + VBMI byte gathers where available; `rec` must then be readable 256 bytes past its end
        expand_obs_records_wide(rec, 0, n, obs_len, obs, 256);
}

int skyjo_host_simd_level(void) { return host_simd_level(); }

int skyjo_host_policy(uint64_t seed, uint64_t genv, uint64_t t, uint32_t legal_bits) {
    if (legal_bits == 0) return -1;
    return policy_pick(seed, genv, t, legal_bits);
}

}  // extern "C"
