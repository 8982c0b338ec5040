// skyjo_hostsimd.h -- host-side expansion of skyjo_step_host's wire format with wide stores
// (csrc/skyjo_hostsimd.cpp, compiled by the host compiler alone: AVX-512 intrinsics never meet nvcc's front end).
#pragma once
#include <stdint.h>

namespace skyjo {

// 0 = portable code only, 2 = AVX-512 (F + BW + VL) paths in use on this CPU, 3 = also AVX-512 VBMI
int host_simd_level();

// Packed words [e0, e1) (bits 0..25 legal actions, 26..27 done, 28..31 agent) -> mask int8[.,26] / agent / done.
// With AVX-512 and 64-byte aligned destinations, full groups of 64 envs are written with non-temporal stores
// (no read-for-ownership of lines that are overwritten completely); the rest goes through the scalar code.
// Ends with a store fence.  All three outputs must be non-null.
void expand_packed_wide(const uint32_t *packed, long long e0, long long e1, int8_t *mask, int8_t *agent, uint8_t *done);

// Compact observation records [e0, e1) -> rows of D bytes, written with non-temporal stores when AVX-512 is
// available and the destination is 64-byte aligned: byte gathers straight from the records with VBMI (needs
// rec_slack >= 256 readable bytes behind record e1 - 1), else staged in L1 per 64 envs.
void expand_obs_records_wide(const uint8_t *rec, long long e0, long long e1, int D, int8_t *obs, long long rec_slack = 0);

}  // namespace skyjo
