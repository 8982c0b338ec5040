// skyjo_policy.h -- launch interface of the fused policy kernel (csrc/skyjo_policy.cu)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace skyjo {

constexpr int POLICY_TILE = 128;     // envs per MMA tile = TMEM lanes
constexpr int POLICY_HIDDEN = 256;   // RLlib TorchFC fcnet_hiddens = [256, 256]
constexpr int POLICY_K1 = 96;        // observation length padded to a multiple of the MMA K (16): D <= 96
constexpr int POLICY_MAX_OBS = 96;
constexpr int POLICY_PACKED_BYTES = POLICY_K1 * POLICY_HIDDEN * 2 + POLICY_HIDDEN * POLICY_HIDDEN * 2 +
                                    POLICY_HIDDEN * 32 * 2 + 2 * POLICY_HIDDEN * 4 + 32 * 4;

struct PolicyParams {
    const int8_t *obs;       // [B, D]
    const int8_t *mask;      // [B, 26]
    const uint8_t *packed;   // POLICY_PACKED_BYTES, written by launch_policy_pack
    long long B;
    int D;
    int bulk_ok;             // obs / mask 16-byte aligned
    unsigned long long first_env, seed, t;
    uint8_t *actions;        // [B] or null
    float *logp, *entropy;   // [B] or null
    float *logits;           // [B, 26] or null (unmasked logits)
    float *value;            // [B]: when set, the packed network is a value head and nothing else is written
    float *dbg1, *dbg2;      // [B, 256] pre-activations of the two hidden layers (tests), or null
    long long *trace;        // [3][POLICY_TRACE_LEN] clock64() marks of CTA 0's three roles (measurement aid), or null
};
constexpr int POLICY_TRACE_LEN = 2048;

cudaError_t launch_policy_pack(const float *w1, const float *b1, const float *w2, const float *b2, const float *w3,
                               const float *b3, int D, int n_out, void *packed, cudaStream_t s);
cudaError_t launch_policy(const PolicyParams &p, int sm_count, cudaStream_t s);

}  // namespace skyjo
