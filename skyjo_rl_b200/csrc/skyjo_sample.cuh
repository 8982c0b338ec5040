// skyjo_sample.cuh -- fused masked softmax + categorical sample for the policy side of a rollout
// (BASELINE config 4).  Replaces, for every env in one launch, what RLlib's TorchCategorical does
// on the output of TorchActionMaskModel (reference rlskyjo/models/action_mask_model.py:63-71:
// logits + clamp(log(mask), FLOAT_MIN), then softmax / sample / logp): illegal actions get
// probability zero exactly, the action is drawn by inverse-CDF from one Philox uniform keyed by
// (sample seed, global env, lockstep t), and the log-probability of the drawn action is returned
// for the PPO ratio.  One env per thread; a row of logits is 104 B (13 x 8-byte loads), a mask
// row 26 B (13 x 2-byte loads); the uint8 actions go straight into skyjo_step.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "skyjo_rng.cuh"

namespace skyjo {

enum : uint32_t { PURPOSE_SAMPLE = 5 };

// Masked softmax + inverse-CDF sample of one action from 26 logits (l is overwritten with the unnormalised
// weights).  `legal` != 0: bit a = action a is legal.  Shared by sample_actions_kernel and the fused policy
// kernel (skyjo_policy.cu), so that equal logits draw equal actions.
__device__ __forceinline__ void sample_masked(float (&l)[26], uint32_t legal, unsigned long long seed,
                                              unsigned long long genv, unsigned long long t, int &act, float &logp,
                                              float &entropy) {
    float mx = -3.0e38f;
#pragma unroll
    for (int a = 0; a < 26; ++a)
        if ((legal >> a) & 1u) mx = fmaxf(mx, l[a]);
    float sum = 0.f, wsum = 0.f;  // sum of exp(l - mx), sum of exp(l - mx) * (l - mx) over legal actions
#pragma unroll
    for (int a = 0; a < 26; ++a) {
        const float d = l[a] - mx;
        const float w = ((legal >> a) & 1u) ? expf(d) : 0.f;
        l[a] = w;
        sum += w;
        wsum += w * (((legal >> a) & 1u) ? d : 0.f);
    }
    const U4 r = rng_block(seed, genv, PURPOSE_SAMPLE, (uint32_t)t, (uint32_t)(t >> 32));
    const float u = (float)(r.x >> 8) * (1.0f / 16777216.0f);  // [0, 1) with 24 bits
    const float target = u * sum;
    act = 31 - __clz((int)legal);  // last legal action: the fallback when rounding leaves cum <= target
    float cum = 0.f;
    bool found = false;
#pragma unroll
    for (int a = 0; a < 26; ++a) {
        cum += l[a];
        if (!found && ((legal >> a) & 1u) && cum > target) {
            act = a;
            found = true;
        }
    }
    float wa = 0.f;
#pragma unroll
    for (int a = 0; a < 26; ++a)
        if (a == act) wa = l[a];
    const float logz = logf(sum);
    logp = logf(wa) - logz;
    entropy = logz - wsum / sum;  // -sum p log p
}

#ifndef SKYJO_SAMPLE_DEVICE_ONLY  // skyjo_policy.cu only wants sample_masked
__global__ void __launch_bounds__(128) sample_actions_kernel(const float *__restrict__ logits,
                                                             const int8_t *__restrict__ mask, long long B,
                                                             unsigned long long first_env, unsigned long long seed,
                                                             unsigned long long t, uint8_t *__restrict__ actions,
                                                             float *__restrict__ logp, float *__restrict__ entropy) {
    const long long e = (long long)blockIdx.x * 128 + threadIdx.x;
    if (e >= B) return;
    float l[26];
    const float2 *lr = reinterpret_cast<const float2 *>(logits + e * 26);
#pragma unroll
    for (int k = 0; k < 13; ++k) {
        const float2 v = lr[k];
        l[2 * k] = v.x;
        l[2 * k + 1] = v.y;
    }
    uint32_t legal = 0;
    const uint16_t *mr = reinterpret_cast<const uint16_t *>(mask + e * 26);
#pragma unroll
    for (int k = 0; k < 13; ++k) {
        const uint32_t v = mr[k];
        legal |= ((v & 0xFFu) ? 1u : 0u) << (2 * k);
        legal |= ((v >> 8) ? 1u : 0u) << (2 * k + 1);
    }
    if (legal == 0) {  // cannot happen for a live env (the mask is never empty); keep the step well defined
        actions[e] = 255;
        logp[e] = 0.f;
        if (entropy) entropy[e] = 0.f;
        return;
    }
    int act;
    float lp, ent;
    sample_masked(l, legal, seed, first_env + (unsigned long long)e, t, act, lp, ent);
    actions[e] = (uint8_t)act;
    logp[e] = lp;
    if (entropy) entropy[e] = ent;
}
#endif

}  // namespace skyjo
