/*
 * c_host.c -- the C ABI of libskyjo_b200.so from a plain C host: no Python, no torch.
 *
 * What a non-Python consumer of the reference's hot path would write: device buffers from cudaMalloc, the env
 * created / bound / reset through include/skyjo_b200.h, the loop of rlskyjo/game/sample_game.py:10-21 for a whole
 * batch with the in-kernel uniform legal policy (skyjo_step_random), then a few steps with EXTERNAL actions chosen on
 * the host from the action mask (skyjo_step) and one step through the host-buffer entry (skyjo_step_host).
 *
 *   gcc -std=c99 -Iinclude -I/usr/local/cuda/include examples/c_host.c -o c_host \
 *       -Lskyjo_rl_b200 -lskyjo_b200 -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/skyjo_rl_b200
 *
 * Built and run by tests/test_c_host.py (compile + link on every CPU run, execution under -m gpu).
 */
#include <cuda_runtime_api.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "skyjo_b200.h"

#define CHECK(call)                                                                      \
    do {                                                                                 \
        int rc_ = (call);                                                                \
        if (rc_ != 0) {                                                                  \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc_, skyjo_last_error());         \
            return 1;                                                                    \
        }                                                                                \
    } while (0)
#define CUDA(call)                                                                       \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess) {                                                         \
            fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e_));                \
            return 1;                                                                    \
        }                                                                                \
    } while (0)

int main(void) {
    const int64_t B = 4096;
    SkyjoConfig cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.num_players = 4; /* BASELINE config 2 */
    cfg.observe_other_player_indirect = 0;
    cfg.score_penalty = 2.0;
    cfg.mean_reward = 1.0;
    cfg.reward_refunded = 0.0;
    cfg.auto_reset = SKYJO_RESET_SAME_STEP;
    const int N = cfg.num_players, D = skyjo_obs_len(&cfg);
    if (skyjo_abi_version() != SKYJO_ABI_VERSION || D != 19 + 12 * N) return 1;

    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
        /* no GPU: the library must say so, loudly (there is no CPU fallback) */
        SkyjoHandle *none = NULL;
        const int rc = skyjo_create(&cfg, 0, B, 1, 0, NULL, 0, &none);
        printf("no CUDA device: skyjo_create -> %d (%s)\n", rc, skyjo_last_error());
        return rc == SKYJO_E_NO_DEVICE ? 0 : 1;
    }

    void *state = NULL;
    SkyjoOutputs out;
    const int64_t state_bytes = skyjo_state_bytes(&cfg, B);
    CUDA(cudaMalloc(&state, (size_t)state_bytes));
    CUDA(cudaMalloc(&out.obs_dev, (size_t)(B * D)));
    CUDA(cudaMalloc(&out.action_mask_dev, (size_t)(B * SKYJO_NUM_ACTIONS)));
    CUDA(cudaMalloc(&out.agent_dev, (size_t)B));
    CUDA(cudaMalloc(&out.done_dev, (size_t)B));
    CUDA(cudaMalloc(&out.reward_dev, (size_t)(B * N) * sizeof(double)));
    CUDA(cudaMalloc(&out.final_score_dev, (size_t)(B * N) * sizeof(double)));

    SkyjoHandle *h = NULL;
    CHECK(skyjo_create(&cfg, 0, B, /*seed*/ 7, /*first_global_env_id*/ 0, state, state_bytes, &h));
    CHECK(skyjo_bind_outputs(h, &out));
    CHECK(skyjo_reset(h, 0));

    /* sample_game.py:10-21 for every env, 300 lockstep steps with the uniform legal policy drawn in-kernel */
    CHECK(skyjo_step_random(h, 300, 0));
    int64_t stats[SKYJO_NUM_STATS];
    CHECK(skyjo_stats_host(h, stats, 0));
    printf("step_random: %lld env-steps, %lld finished games, mean length %.1f, %lld illegal\n",
           (long long)stats[SKYJO_STAT_STEPS], (long long)stats[SKYJO_STAT_EPISODES],
           (double)stats[SKYJO_STAT_EPISODE_STEPS] / (double)(stats[SKYJO_STAT_EPISODES] ? stats[SKYJO_STAT_EPISODES] : 1),
           (long long)stats[SKYJO_STAT_ILLEGAL]);
    if (stats[SKYJO_STAT_STEPS] != 300 * B || stats[SKYJO_STAT_EPISODES] < B || stats[SKYJO_STAT_ILLEGAL] != 0) return 1;

    /* external actions: the last legal action of every env's mask, chosen on the host */
    int8_t *mask = (int8_t *)malloc((size_t)(B * SKYJO_NUM_ACTIONS));
    uint8_t *act = (uint8_t *)malloc((size_t)B), *act_dev = NULL;
    CUDA(cudaMalloc((void **)&act_dev, (size_t)B));
    for (int t = 0; t < 20; ++t) {
        CUDA(cudaMemcpy(mask, out.action_mask_dev, (size_t)(B * SKYJO_NUM_ACTIONS), cudaMemcpyDeviceToHost));
        for (int64_t e = 0; e < B; ++e) {
            int a = -1;
            for (int k = 0; k < SKYJO_NUM_ACTIONS; ++k)
                if (mask[e * SKYJO_NUM_ACTIONS + k]) a = k;
            if (a < 0) return 1; /* the mask is never empty */
            act[e] = (uint8_t)a;
        }
        CUDA(cudaMemcpy(act_dev, act, (size_t)B, cudaMemcpyHostToDevice));
        CHECK(skyjo_step(h, act_dev, SKYJO_ACT_U8, 0));
    }
    /* and one step through the host-buffer entry: actions in, observation / mask / agent / done / reward out */
    int8_t *obs_h = (int8_t *)malloc((size_t)(B * D)), *agent_h = (int8_t *)malloc((size_t)B);
    uint8_t *done_h = (uint8_t *)malloc((size_t)B);
    double *rew_h = (double *)malloc((size_t)(B * N) * sizeof(double));
    CUDA(cudaMemcpy(mask, out.action_mask_dev, (size_t)(B * SKYJO_NUM_ACTIONS), cudaMemcpyDeviceToHost));
    for (int64_t e = 0; e < B; ++e)
        for (int k = 0; k < SKYJO_NUM_ACTIONS; ++k)
            if (mask[e * SKYJO_NUM_ACTIONS + k]) act[e] = (uint8_t)k;
    CHECK(skyjo_step_host(h, act, obs_h, mask, agent_h, done_h, rew_h, 0));
    int8_t *obs_chk = (int8_t *)malloc((size_t)(B * D));
    CUDA(cudaMemcpy(obs_chk, out.obs_dev, (size_t)(B * D), cudaMemcpyDeviceToHost));
    if (memcmp(obs_chk, obs_h, (size_t)(B * D)) != 0) return 1;
    CHECK(skyjo_check(h, 0));
    CHECK(skyjo_stats_host(h, stats, 0));
    printf("external actions: %lld env-steps in total, %lld illegal; host entry returned agent %d, obs[0..2] = %d %d %d\n",
           (long long)stats[SKYJO_STAT_STEPS], (long long)stats[SKYJO_STAT_ILLEGAL], (int)agent_h[0], (int)obs_h[0],
           (int)obs_h[1], (int)obs_h[2]);
    if (stats[SKYJO_STAT_STEPS] != 321 * B || stats[SKYJO_STAT_ILLEGAL] != 0) return 1;
    CHECK(skyjo_destroy(h));
    puts("c_host ok");
    return 0;
}
