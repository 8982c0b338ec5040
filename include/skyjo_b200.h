/*
 * skyjo_b200.h -- C ABI of libskyjo_b200.so, the B200-native batched SkyJo environment.
 *
 * The reference (michaelfeil/skyjo_rl) has no FFI: its boundary is two Python classes,
 * `SkyjoGame` (rlskyjo/game/skyjo.py:19) and the PettingZoo env `SimpleSkyjoEnv`
 * (rlskyjo/environment/skyjo_env.py:29).  This header is the boundary a replacement
 * binds instead; each entry point names the reference interface it replaces.  The Python
 * host (skyjo_rl_b200/env.py) loads it with ctypes; INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - plain C types only; every *_dev pointer is a CUDA device pointer owned by the caller
 *     (torch tensors' data_ptr()), every *_host pointer is host memory;
 *   - `stream` is a cudaStream_t passed as void* (0 = default stream); calls are
 *     asynchronous on it unless the comment says "synchronises";
 *   - return value 0 = success, otherwise an SKYJO_E_* code (or 1000 + cudaError_t);
 *     the message is available from skyjo_last_error() (thread-local);
 *   - illegal actions are data, not errors: the env terminates with reward -1 for the
 *     offender (TerminateIllegalWrapper semantics, skyjo_env.py:23).
 *   - there is no CPU fallback: without a CUDA device every compute entry fails.
 */
#ifndef SKYJO_B200_H
#define SKYJO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SKYJO_ABI_VERSION 1
#define SKYJO_MAX_PLAYERS 12
#define SKYJO_NUM_ACTIONS 26 /* skyjo.py:46 action_mask_shape */
#define SKYJO_DECK 150       /* skyjo.py:80 */
#define SKYJO_NUM_STATS 32

enum {
    SKYJO_OK = 0,
    SKYJO_E_INVALID = 1,    /* bad argument / config (skyjo.py:24-26 assert) */
    SKYJO_E_NOT_BOUND = 2,  /* outputs not bound */
    SKYJO_E_STATE = 3,      /* device-side consistency flag raised (see skyjo_check) */
    SKYJO_E_NO_DEVICE = 4,
    SKYJO_E_NCCL = 5,       /* ncclAllReduce returned an error (skyjo_stats_allreduce) */
    SKYJO_E_CUDA = 1000     /* + cudaError_t */
};

/* action tensor element types accepted by skyjo_step */
enum { SKYJO_ACT_U8 = 0, SKYJO_ACT_I8 = 1, SKYJO_ACT_I32 = 2, SKYJO_ACT_I64 = 3 };

/* values of the per-env `done` byte written by every step */
enum {
    SKYJO_RUNNING = 0,
    SKYJO_DONE_GAME_OVER = 1, /* skyjo.py:350-356 */
    SKYJO_DONE_ILLEGAL = 2,   /* skyjo_env.py:23 */
    SKYJO_DONE_TRUNCATED = 3  /* step cap (vanilla_env_example.py:14 uses 300*N) */
};

/* indices into the statistics vector (int64[SKYJO_NUM_STATS]) */
enum {
    SKYJO_STAT_EPISODES = 0,          /* games that reached game over */
    SKYJO_STAT_EPISODE_STEPS = 1,     /* sum of their lengths in act() calls */
    SKYJO_STAT_SCORE_RAW_SUM = 2,     /* sum over seats of the unpenalised score */
    SKYJO_STAT_WINNER_RAW_SUM = 3,    /* sum of min-over-seats raw score */
    SKYJO_STAT_FINISHER_RAW_SUM = 4,  /* sum of the finisher's raw score */
    SKYJO_STAT_PENALISED = 5,         /* games whose finisher was penalised (skyjo.py:496) */
    SKYJO_STAT_PENALISED_RAW_SUM = 6, /* sum of those finishers' raw scores */
    SKYJO_STAT_REFUNDS = 7,           /* column removals in finished games (skyjo.py:419) */
    SKYJO_STAT_RESHUFFLES = 8,        /* in-game draw-pile reshuffles (skyjo.py:361-365) */
    SKYJO_STAT_ILLEGAL = 9,
    SKYJO_STAT_TRUNCATED = 10,
    SKYJO_STAT_ACT_DRAW_PILE = 11,    /* action 24 */
    SKYJO_STAT_ACT_TAKE_DISCARD = 12, /* action 25 */
    SKYJO_STAT_ACT_SWAP = 13,         /* actions 0..11 */
    SKYJO_STAT_ACT_FLIP = 14,         /* actions 12..23 */
    SKYJO_STAT_STARTER_SEAT0 = 15,    /* finished games started by seat 0 */
    SKYJO_STAT_STEPS = 16,            /* all act() calls */
    SKYJO_STAT_WINS_SEAT0 = 17        /* .. +11: first argmin of the final scores */
};

/* SkyjoConfig.auto_reset.  The reference leaves reset() to the caller (sample_game.py:8-9,
 * skyjo_env.py:254-267); a batched env has to schedule it per env:
 *   OFF        the env freezes; every further step returns done (skyjo.py:316-321) until skyjo_reset.
 *   SAME_STEP  the step that ends an episode also installs the next one: it publishes the ended
 *              episode's done / rewards / final scores together with the NEW episode's first
 *              observation, mask and agent.
 *   NEXT_STEP  phase-locked: the step that ends an episode publishes done / rewards / final scores
 *              with the TERMINAL observation (agent = the finisher, draw mask, as observe() would
 *              return after game over); the env's next lockstep slot is the reset: its action is
 *              ignored, it is not an env-step (not counted in SKYJO_STAT_STEPS), done = 0, rewards
 *              cleared, and it publishes the new episode's first observation.  An episode is an odd
 *              number of act() calls, so all envs of the batch stay in the same phase (draw on even,
 *              place on odd slots) and a warp never executes both transitions.  An episode that
 *              ends in a place slot (illegal placement, truncation after a placement) is replaced
 *              at once, which keeps the lock. */
enum { SKYJO_RESET_OFF = 0, SKYJO_RESET_SAME_STEP = 1, SKYJO_RESET_NEXT_STEP = 2 };

/* Mirrors the keyword arguments of SkyjoGame.__init__ (skyjo.py:20-22) and
 * SimpleSkyjoEnv.__init__ (skyjo_env.py:38-45). */
typedef struct SkyjoConfig {
    int32_t num_players;                   /* 1..12 */
    int32_t observe_other_player_indirect; /* 0: obs has 19+12N entries, 1: 31 */
    double score_penalty;
    double mean_reward;
    double reward_refunded;
    int32_t auto_reset;        /* SKYJO_RESET_*: what happens to an env whose episode has ended */
    int32_t max_episode_steps; /* 0 = no truncation */
} SkyjoConfig;

/* Output buffers (device).  Shapes: obs int8[B, D]; action_mask int8[B, 26];
 * agent int8[B] (= expected_action[0], agent_selection is "player_<agent>");
 * done uint8[B]; reward float64[B, N]; final_score float64[B, N]. */
typedef struct SkyjoOutputs {
    void *obs_dev;
    void *action_mask_dev;
    void *agent_dev;
    void *done_dev;
    void *reward_dev;
    void *final_score_dev;
} SkyjoOutputs;

/* SkyjoGame-shaped view of one env for debugging and parity tests
 * (players_cards / players_masked / hand_card / expected_action / game_metrics). */
typedef struct SkyjoEnvDebug {
    int8_t players_cards[SKYJO_MAX_PLAYERS][12];  /* true values; -14 where refunded */
    int8_t players_masked[SKYJO_MAX_PLAYERS][12]; /* 2 hidden, 1 open, 0 refunded */
    int8_t discard_hist[16];  /* multiset of the discard pile: count of value j-2 in [j] */
    int8_t draw_hist[16];     /* valid when draw_is_multiset */
    int8_t drawpile[SKYJO_DECK]; /* python-list order (top = last), valid otherwise */
    int16_t n_draw;
    int16_t n_discard;
    int8_t hand_card;    /* 15 = none (skyjo.py:33) */
    int8_t discard_top;  /* -3 = empty (skyjo.py:254) */
    int8_t expected_player;
    int8_t expected_phase; /* 0 draw, 1 place */
    int8_t starter;
    int8_t is_terminated;
    int8_t draw_is_multiset; /* 1 after an in-game reshuffle (DESIGN.md "reshuffle") */
    int8_t n_reshuffles;
    int32_t step_in_episode;
    uint32_t episode;
    int8_t num_refunded[SKYJO_MAX_PLAYERS];
    int16_t num_placed[SKYJO_MAX_PLAYERS];
    int8_t pad[8];
} SkyjoEnvDebug;

typedef struct SkyjoHandle SkyjoHandle;

int skyjo_abi_version(void);
const char *skyjo_last_error(void);

/* skyjo.py:43-45 obs_shape */
int skyjo_obs_len(const SkyjoConfig *cfg);
/* bytes of device state the caller must allocate (256-byte aligned) for num_envs games */
int64_t skyjo_state_bytes(const SkyjoConfig *cfg, int64_t num_envs);

/* Replaces SkyjoGame.__init__ / SimpleSkyjoEnv.__init__ for a batch: env i of this handle
 * is global env first_global_env_id + i; its RNG streams depend only on (seed, global id),
 * so any sharding of a global batch over handles/GPUs plays identical games. */
int skyjo_create(const SkyjoConfig *cfg, int device, int64_t num_envs, uint64_t seed,
                 int64_t first_global_env_id, void *state_dev, int64_t state_bytes,
                 SkyjoHandle **out);
int skyjo_destroy(SkyjoHandle *h);
int skyjo_bind_outputs(SkyjoHandle *h, const SkyjoOutputs *outs);

/* SkyjoGame.reset (skyjo.py:52-74) / SimpleSkyjoEnv.reset (skyjo_env.py:254-267) for every
 * env: Philox deck shuffle, deal, two flips per player, start player; writes obs / mask /
 * agent for the first turn, zeroes done / reward. */
int skyjo_reset(SkyjoHandle *h, void *stream);
/* Same with the shuffles and flips injected: decks int8[B,150], flips
 * uint8[B,N,2].  Decks must hold values -2..12, at most 15 copies of each. */
int skyjo_reset_injected(SkyjoHandle *h, const int8_t *decks_dev, const uint8_t *flips_dev,
                         void *stream);
/* SimpleSkyjoEnv.seed (skyjo_env.py:280-290): new seed, episode counters to 0, then reset */
int skyjo_seed(SkyjoHandle *h, uint64_t seed, void *stream);

/* One fused launch per call = SimpleSkyjoEnv.step (skyjo_env.py:216-252) + observe
 * (:199-214) for the next agent, for all envs: SkyjoGame.act (skyjo.py:308-335),
 * _calc_final_rewards (skyjo_env.py:293-312), collect_observation (skyjo.py:148-199). */
int skyjo_step(SkyjoHandle *h, const void *actions_dev, int action_dtype, void *stream);
/* n_steps fused launches with the uniform legal policy (random_admissible_policy.py:26-28)
 * drawn in-kernel: the loop of sample_game.py:10-21.  From the second call with the same n_steps
 * (>= 8) the whole launch sequence of a call -- step kernels of every env range, refill deals, the
 * fork / join between the range streams -- is replayed from an instantiated CUDA graph whose kernels
 * read the lockstep counter from device memory: one cudaGraphLaunch instead of ~4 n kernel launches
 * (0.025 instead of 0.97 ms of host time per 64 steps at 2^20 envs).  Same games bit for bit;
 * env SKYJO_NO_GRAPH=1 keeps the direct launches.  Batches of one range use the graph up to 2^15 envs. */
int skyjo_step_random(SkyjoHandle *h, int n_steps, void *stream);
/* skyjo_step_random steps a batch as n independent env ranges, each on its own CUDA stream with its own
 * refill deals (games never interact; one range's launch ramp / tail is covered by the others' CTAs).
 * n = 0 restores the default (4 ranges from 2^18 envs, else 1; env SKYJO_RANGES overrides at creation),
 * n = 1..8 forces it -- results are bit-identical for every n (tested against the oracle). */
int skyjo_set_env_ranges(SkyjoHandle *h, int n);
/* Time-major rollout buffers (device) of skyjo_rollout_random: obs int8[T, B, D];
 * action_mask int8[T, B, 26]; agent int8[T, B]; done uint8[T, B]. */
typedef struct SkyjoRollout {
    void *obs_dev;
    void *action_mask_dev;
    void *agent_dev;
    void *done_dev;
} SkyjoRollout;
/* The same n_steps env-steps as skyjo_step_random (identical games, statistics and final state),
 * run as multi-step launches: each kernel advances its envs by up to 32 consecutive env-steps
 * (the refill window: 8 / 16 / 24 / 32 for 1 / 2 / 3 / >= 4 players) with
 * the state held in registers and stores what step t would have published -- the next agent's
 * observation, action mask, agent and done code -- into slice t of the time-major buffers (the
 * rollout storage a learner consumes, sample_game.py:10-21 unrolled in time).  The bound [B, ...]
 * outputs receive the last step's values; reward / final_score keep their [B, N] "last finished
 * episode" meaning (an episode that ends inside a launch is visible in the statistics, its
 * reward row is cleared by the env's next step as in skyjo_env.py:250-252). */
int skyjo_rollout_random(SkyjoHandle *h, int n_steps, const SkyjoRollout *out, void *stream);
/* CUDA-event timing of everything launched between begin and end on `stream` (summed device ms
 * and launch counts of step / rollout kernels and of deal kernels); end synchronises.  One event
 * pair brackets each deal or rollout launch, and each WINDOW of back-to-back single-step launches
 * (the up to 32 launches between two refill deals): an event between two ~45 us step kernels would
 * add ~5 us to each pair and break their programmatic dependent launch.  Measurement aid for
 * bench.py's roofline figure. */
int skyjo_profile_begin(SkyjoHandle *h);
int skyjo_profile_end(SkyjoHandle *h, void *stream, double *step_ms, double *deal_ms,
                      int64_t *n_step_launches, int64_t *n_deal_launches);
/* skyjo_step_random on one stream between skyjo_profile_begin / _end: returns the summed
 * device time (ms) and launch counts of the step kernels and of the deal kernels; synchronises.
 * Measurement aid for bench.py's roofline figure. */
int skyjo_step_random_profile(SkyjoHandle *h, int n_steps, void *stream, double *step_ms,
                              double *deal_ms, int64_t *n_step_launches, int64_t *n_deal_launches);
/* Host-buffer entry for end-to-end use (env.step(a); env.last() of vanilla_env_example.py:14-35
 * for every env): copies actions (uint8[B]) to the device, steps, and fills the host buffers
 * obs int8[B,D] / mask int8[B,26] / agent int8[B] / done uint8[B] / reward float64[B,N] with
 * exactly what the device output buffers hold.  Null outputs are skipped.  Returns when the host
 * buffers are complete (the refill deal of finished envs may still be running on `stream`).
 * Wire format (csrc/skyjo_hostio.cuh): mask + agent + done cross the link as one 32-bit word per
 * env and are expanded on the host by a few worker threads; reward rows cross only for envs
 * whose episode ended in this step, the other rows are zero-filled on the host; observation
 * rows are copied as they are (or as compact records, skyjo_set_host_wire).  Use pinned host
 * buffers for full link speed, and the same reward buffer on consecutive calls.
 * Batches of at most 256 envs take a latency path instead: the actions are read through a host-mapped pointer,
 * one block writes all outputs into host-mapped pinned memory, the host polls a sequence word and copies them
 * into the caller's buffers (any host memory), and the refill deal is launched only when an episode ended. */
int skyjo_step_host(SkyjoHandle *h, const uint8_t *actions_host, int8_t *obs_host,
                    int8_t *mask_host, int8_t *agent_host, uint8_t *done_host,
                    double *reward_host, void *stream);
/* worker threads skyjo_step_host uses for the host-side expansion; 0 = default
 * (this rank's share of the CPUs the process may run on -- slice LOCAL_RANK of LOCAL_WORLD_SIZE under torchrun --
 * minus one, between 2 and 8, or SKYJO_HOST_THREADS).  The workers are persistent and pinned one per core of that slice
 * (SKYJO_HOST_PIN=0 disables pinning). */
int skyjo_set_host_threads(SkyjoHandle *h, int n);
/* how observation rows cross the link in skyjo_step_host, per env range of the batch: 0 = as they are, straight
 * into obs_host by the copy engine (no CPU work); 1 = compact records of 12 + 6 R + ceil(R / 2) bytes for a row of
 * 19 + 12 R bytes, packed on the device and expanded on the host by the worker threads (1.7x fewer bytes on the
 * link, paid for in host memory traffic); 2 = mixed: the first k of 8 ranges compact, the rest raw, k tuned between
 * calls by measurement (perturb and observe on the call time, starting from k = 4 / 2 / 0 for 1 / 2-3 / >= 4 ranks
 * on the host, LOCAL_WORLD_SIZE).  Default: 2 where the CPU has the
 * AVX-512 streaming expansions (skyjo_host_simd_level() == 2), else 0; env SKYJO_HOST_WIRE=raw|compact|mixed
 * selects at creation, SKYJO_HOST_MIX=k pins the share.  All modes fill the host buffers bit-identically. */
int skyjo_set_host_wire(SkyjoHandle *h, int mode);
/* bytes the last skyjo_step_host call moved device -> host */
int64_t skyjo_host_wire_bytes(const SkyjoHandle *h);
/* of 8 env ranges, how many the next skyjo_step_host call sends as compact records (0 raw .. 8 compact) */
int skyjo_host_wire_share(const SkyjoHandle *h);

/* SimpleSkyjoEnv.observe(agent) (skyjo_env.py:199-214) for an arbitrary seat; agent = -1
 * means each env's agent_selection.  Writes int8[B,D] / int8[B,26]. */
int skyjo_observe(SkyjoHandle *h, int agent, void *obs_dev, void *mask_dev, void *stream);

/* Episode statistics (SkyjoGame.game_metrics aggregated, skyjo.py:56-60).  *_device reduces
 * into int64[SKYJO_NUM_STATS] on the device (feed it to an NCCL all-reduce); *_host also
 * copies to the host and synchronises. */
int skyjo_stats_device(SkyjoHandle *h, int64_t *out_dev, void *stream);
int skyjo_stats_host(SkyjoHandle *h, int64_t *out_host, void *stream);
/* The library's only collective (never on the step path): skyjo_stats_device into out_dev followed
 * by ncclAllReduce(sum, int64) over the ranks of `nccl_comm` (an ncclComm_t) on `stream`, in place.
 * NCCL is resolved at run time from the libnccl already loaded in the process (torch's, or one the
 * caller opened), never linked: fails with SKYJO_E_INVALID if there is none. */
int skyjo_stats_allreduce(SkyjoHandle *h, void *nccl_comm, int64_t *out_dev, void *stream);
/* The same collective off the caller's stream: `stream` only runs the one-CTA device-side reduction
 * (a snapshot of the counters at that point); the ncclAllReduce into out_dev is queued on a stream the library
 * owns, so launches queued on `stream` afterwards never wait for a peer rank.  out_dev is valid once
 * skyjo_stats_allreduce_wait has made a stream wait for it (or after skyjo_destroy); keep it alive and unread
 * until then, and alternate between two buffers when issuing a call per iteration.  nccl_comm == NULL sums over
 * one rank (a device copy), so single-GPU loops have the same shape. */
int skyjo_stats_allreduce_async(SkyjoHandle *h, void *nccl_comm, int64_t *out_dev, void *stream);
int skyjo_stats_allreduce_wait(SkyjoHandle *h, void *stream);
int skyjo_stats_clear(SkyjoHandle *h, void *stream);

/* Policy side of a rollout (BASELINE config 4): masked softmax + categorical sample of one action
 * per env in one launch -- what RLlib's TorchCategorical does on TorchActionMaskModel's output
 * (action_mask_model.py:63-71).  logits float32[B,26]; mask int8[B,26] (null = the bound
 * action_mask output); illegal actions have probability exactly 0; the draw is an inverse-CDF
 * sample from one Philox uniform keyed by (sample_seed, global env id, lockstep counter).
 * Writes actions uint8[B] (feed them to skyjo_step), logp float32[B] of the drawn action and,
 * if not null, the entropy float32[B] of the masked distribution. */
int skyjo_sample_actions(SkyjoHandle *h, const float *logits_dev, const int8_t *mask_dev,
                         uint64_t sample_seed, uint8_t *actions_dev, float *logp_dev,
                         float *entropy_dev, void *stream);
/* The whole policy side of a rollout step in ONE kernel (csrc/skyjo_policy.cu): the reference's action-mask model
 * (action_mask_model.py:58-77: RLlib TorchFC, two tanh layers of 256, logits + clamp(log(mask), FLOAT_MIN)) evaluated
 * on the bound int8 observation rows in place -- bf16 operands, fp32 accumulation, tcgen05.mma with the three weight
 * matrices resident in shared memory and every activation in tensor memory -- then the masked softmax + categorical
 * sample of skyjo_sample_actions (same Philox keying: equal logits draw equal actions).  Observation rows of at most
 * 96 bytes (direct mode up to 6 players, or indirect mode); functional (bf16) parity with an fp32 evaluation, the
 * tolerance is stated in tests/test_gpu_fused_policy.py.
 *   skyjo_policy_pack: float32 torch.nn.Linear parameters (w[out][in] row-major, device pointers: w1 [256, obs_len],
 *     b1 [256], w2 [256, 256], b2 [256], w3 [n_out, 256], b3 [n_out]; n_out = 26 for the logits net, 1 for a value
 *     head) -> the kernel's weight image (skyjo_policy_packed_bytes() bytes of device memory, 16-byte aligned).
 *   skyjo_policy_sample: actions uint8[B] (feed them to skyjo_step), and if not null logp float32[B] of the drawn
 *     action, entropy float32[B] of the masked distribution, the unmasked logits float32[B, 26].
 *   skyjo_policy_value: value float32[B] from a packed value head.
 *   skyjo_policy_debug: logits and the pre-activations float32[B, 256] of the two hidden layers (parity tests). */
int64_t skyjo_policy_packed_bytes(void);
int skyjo_policy_pack(int obs_len, int n_out, const float *w1_dev, const float *b1_dev, const float *w2_dev,
                      const float *b2_dev, const float *w3_dev, const float *b3_dev, void *packed_dev, void *stream);
int skyjo_policy_sample(SkyjoHandle *h, const void *packed_dev, uint64_t sample_seed, uint8_t *actions_dev,
                        float *logp_dev, float *entropy_dev, float *logits_dev, void *stream);
int skyjo_policy_value(SkyjoHandle *h, const void *packed_dev, float *value_dev, void *stream);
int skyjo_policy_debug(SkyjoHandle *h, const void *packed_dev, float *pre1_dev, float *pre2_dev, float *logits_dev,
                       void *stream);
/* Measurement aid: skyjo_policy_sample (seed 0) that also records clock64() at every hand-over between the three
 * roles of the kernel's first CTA into trace_dev, int64[3][skyjo_policy_trace_len()] zero-filled by the caller
 * (hidden-epilogue team, staging / sampling team, MMA warp); decoded by tools/policy_trace.py. */
int skyjo_policy_trace(SkyjoHandle *h, const void *packed_dev, uint8_t *actions_dev, int64_t *trace_dev, void *stream);
int skyjo_policy_trace_len(void);
/* Closes the running refill window and makes `stream` wait for the library's internal streams:
 * work queued on `stream` afterwards may read or overwrite the state buffer (checkpointing). */
int skyjo_quiesce(SkyjoHandle *h, void *stream);
/* SkyjoGame-shaped dump of envs [env0, env0+count) into out_dev (device memory). */
int skyjo_export_debug(SkyjoHandle *h, int64_t env0, int64_t count, SkyjoEnvDebug *out_dev,
                       void *stream);
/* Synchronises; returns SKYJO_E_STATE if a kernel raised the consistency flag. */
int skyjo_check(SkyjoHandle *h, void *stream);
/* lockstep counter (number of step launches since creation / seed) */
int64_t skyjo_step_count(const SkyjoHandle *h);
/* restore the lockstep counter when resuming from a state buffer saved after skyjo_quiesce */
int skyjo_set_step_count(SkyjoHandle *h, int64_t t);
/* kernels launched by this handle so far */
int64_t skyjo_launch_count(const SkyjoHandle *h);
/* calls of skyjo_step_random served by a CUDA-graph replay so far (their kernels are counted by
 * skyjo_launch_count like directly launched ones) */
int64_t skyjo_graph_replay_count(const SkyjoHandle *h);

/* Host twins of the device RNG streams, so a CPU checker can be dealt identical games. */
void skyjo_host_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void skyjo_host_deck(uint64_t seed, uint64_t global_env, uint32_t episode, int8_t out[SKYJO_DECK]);
void skyjo_host_flips(uint64_t seed, uint64_t global_env, uint32_t episode, int num_players,
                      uint8_t *out /* [N][2] */);
int skyjo_host_policy(uint64_t seed, uint64_t global_env, uint64_t t, uint32_t legal_bits);
/* Host twin of the in-game draw-pile reshuffle, i.e. the rule that stands in for
 * SkyjoGame._reshuffle_discard_pile (skyjo.py:127-138, called from :361-365): permutes
 * pile[0..len) (card values -2..12, len <= 512) in place into the order the device plays for the
 * reshuffle_index-th reshuffle (0-based) of `episode` of `global_env`.  Python-list convention of
 * the reference: the LAST element becomes the new discard card (`drawpile.pop()`, :137), the one
 * before it is the first card drawn, and so on down to pile[0].  The order is sequential sampling
 * without replacement from the pile's multiset with Philox numbers -- a uniform random permutation.
 * A third party that drives the unmodified reference with
 *     SkyjoGame._reshuffle_discard_pile = staticmethod(f)   # f calls this, see INTEGRATION.md
 * gets the games the device plays.  Returns 0, or SKYJO_E_INVALID for a bad pile. */
int skyjo_host_reshuffle(uint64_t seed, uint64_t global_env, uint32_t episode,
                         uint32_t reshuffle_index, int8_t *pile, int len);
/* Host half of skyjo_step_host's wire format: expands n packed words (bits 0..25 legal actions,
 * 26..27 done code, 28..31 agent) into mask int8[n,26] / agent int8[n] / done uint8[n]
 * (null outputs are skipped). */
void skyjo_host_expand_packed(const uint32_t *packed, int64_t n, int8_t *mask, int8_t *agent,
                              uint8_t *done);
/* The compact observation record of skyjo_step_host (csrc/skyjo_hostio.cuh): bytes per record for
 * rows of obs_len bytes (-1 if obs_len is not 19 + 12 R); the host twin of the device-side
 * packing (returns the number of rows that are not encodable); and the host-side expansion
 * (portable: 0 = the fastest version this CPU has -- `rec` must then be readable 256 bytes past its end --, 1 = scalar,
 * 2 = 16-byte shuffles with plain stores, 3 = those staged in L1 and streamed out). */
int skyjo_host_obs_record_bytes(int obs_len);
/* 2 when the host-side expansions run their AVX-512 / streaming-store versions on this CPU, 3 with AVX-512 VBMI
 * (record expansion by byte gathers), else 0 */
int skyjo_host_simd_level(void);
int64_t skyjo_host_pack_obs(const int8_t *obs, int64_t n, int obs_len, uint8_t *rec);
void skyjo_host_expand_obs(const uint8_t *rec, int64_t n, int obs_len, int8_t *obs, int portable);

#ifdef __cplusplus
}
#endif
#endif /* SKYJO_B200_H */
