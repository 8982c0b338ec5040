// skyjo_encode.cuh -- device-side staging of the observation / action-mask rows.
//
// A CTA owns TILE consecutive envs, so its slice of the row-major outputs obs[B, D] /
// mask[B, 26] is one contiguous span.  Each thread turns its env's row into a word stream
// (encode_words, skyjo_core.cuh), shifts it to the row's byte offset in the tile -- rows are
// 67 / 115 / 31 / 26 bytes, never a multiple of 4 -- taking the few bytes its last word shares
// with the next row from the neighbouring lane by a warp shuffle, and writes only aligned
// 32-bit words to shared memory.  The finished span leaves with one TMA bulk store
// (cp.async.bulk.global.shared::cta) per tensor.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "skyjo_core.cuh"
#include "skyjo_state.cuh"

namespace skyjo {

template <int NW, int TAIL>
__device__ __forceinline__ void stage_stream(const uint32_t (&S)[NW], uint32_t *s_tile, int tid) {
    const uint32_t next0 = __shfl_down_sync(0xFFFFFFFFu, S[0], 1);
    uint32_t out[NW], first;
    int count;
    stage_words<NW, TAIL>(S, next0, tid, out, first, count);
    uint32_t *dst = s_tile + first;
#pragma unroll
    for (int k = 0; k < NW - 1; ++k) dst[k] = out[k];
    if (count == NW) dst[NW - 1] = out[NW - 1];
}

// Copies the CTA's staged tile to global memory.  Full tiles with 16-byte aligned targets go
// out as one TMA bulk store issued by thread 0; ragged tails fall back to a byte loop.
__device__ __forceinline__ void store_tile(const uint8_t *s_src, int8_t *g_dst, uint32_t nbytes, bool bulk,
                                           int tid) {
    if (bulk) {
        if (tid == 0) {
            uint32_t saddr = (uint32_t)__cvta_generic_to_shared(s_src);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(g_dst),
                         "r"(saddr), "r"(nbytes)
                         : "memory");
        }
    } else {
        for (uint32_t i = tid; i < nbytes; i += TILE) g_dst[i] = (int8_t)s_src[i];
    }
}

__device__ __forceinline__ void bulk_commit_and_wait(int tid) {
    if (tid == 0) {
        asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
    }
}

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

}  // namespace skyjo
