// skyjo_policy.cu -- the consumer side of BASELINE config 4 as ONE sm_100a kernel: the action-mask policy of the
// reference (rlskyjo/models/action_mask_model.py:58-77 -- RLlib's TorchFC with fcnet_hiddens [256, 256], tanh,
// logits + clamp(log(mask), FLOAT_MIN)) evaluated on the env's int8 observation rows in place, followed by the
// masked softmax + categorical sample that emits the uint8 actions skyjo_step consumes.
//
//   obs int8[B, D] --bf16--> [x W1^T + b1] -tanh-> [h1 W2^T + b2] -tanh-> [h2 W3^T + b3] -> mask, softmax, sample
//
// The three matrix products are the only contractions near the hot path, and they run on the 5th-generation
// tensor cores: tcgen05.mma (kind::f16, bf16 operands, fp32 accumulation in tensor memory), M = 128 envs per
// tile, issued by one thread.  Layout of one persistent CTA (one per SM, 288 threads):
//   * shared memory holds the three weight matrices for the whole launch (bf16, 192 KB, in the canonical K-major
//     no-swizzle core-matrix layout the MMA descriptors address: 8 rows x 16 bytes per core matrix), the biases,
//     and the obs / mask rows of the tile being started (1-D TMA bulk copies, one tile ahead);
//   * tensor memory (all 512 columns x 128 lanes) holds every activation: an env is a TMEM lane.  The A operand
//     of each product is read FROM TENSOR MEMORY (tcgen05.mma with a TMEM A address), so activations never touch
//     shared memory: one thread per env converts the obs row to bf16 and tcgen05.st's it; after each hidden layer
//     one thread per env reads the fp32 accumulators with tcgen05.ld, adds the bias, applies tanh, packs to bf16 and
//     stores them back in place (a 32-column chunk of fp32 compacts into 16 columns of bf16 pairs), announcing
//     every chunk on an mbarrier so that the next layer's MMAs over that K-chunk start at once;
//   * one team of warps does nothing but the hidden epilogues (the critical path), a second team stages the next
//     tile's operand and samples the previous tile's actions meanwhile, one more warp issues the MMAs (6 + 16 + 16
//     per tile); the two halves of tensor memory swap roles from tile to tile (column map at policy_kernel).
// Roofline: 2 x (96 + 256) x 256 + 2 x 256 x 32 = 196 608 FLOP per env as issued (178 688 useful) on the tensor cores,
// and 512 tanh per env on the SFUs (MUFU.TANH: 16 XU-pipe cycles per warp instruction), which is the longer of the
// two per tile: 57 us per 2^18 envs at 1.965 GHz against 22 us of MMA time; see DESIGN.md.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "skyjo_policy.h"
#define SKYJO_SAMPLE_DEVICE_ONLY
#include "skyjo_sample.cuh"

namespace skyjo {

// Team H has one or two warps per scheduler (with two, warps w and w + 4 share the 32 lanes of quadrant w and take
// one half of the 256 hidden units each).  Measured (2^18 envs, tools/policy_kernel_bench.py): see DESIGN.md.
#ifndef SKYJO_POLICY_HGROUPS
#define SKYJO_POLICY_HGROUPS 1
#endif
constexpr int POL_HGROUPS = SKYJO_POLICY_HGROUPS;
constexpr uint32_t POL_D3_COL = POL_HGROUPS == 1 ? 128u : 64u;  // the logits accumulate in 32 columns of P clear of h1 and of x(j+1)
constexpr int POL_H_WARPS = 4 * POL_HGROUPS;       // warps [0, POL_H_WARPS): hidden epilogues
constexpr int POL_S_WARP0 = POL_H_WARPS;           // 4 warps: operand staging + sampling (one thread per env = TMEM lane)
constexpr int POL_MMA_WARP = POL_H_WARPS + 4;      // MMA issue
constexpr int POL_THREADS = 32 * (POL_MMA_WARP + 1);

constexpr int POL_N3 = 32;        // 26 logits (or 1 value) padded to an MMA N
constexpr uint32_t OFF_W1 = 0;
constexpr uint32_t OFF_W2 = OFF_W1 + POLICY_K1 * POLICY_HIDDEN * 2;
constexpr uint32_t OFF_W3 = OFF_W2 + POLICY_HIDDEN * POLICY_HIDDEN * 2;
constexpr uint32_t OFF_B1 = OFF_W3 + POLICY_HIDDEN * POL_N3 * 2;
constexpr uint32_t OFF_B2 = OFF_B1 + POLICY_HIDDEN * 4;
constexpr uint32_t OFF_B3 = OFF_B2 + POLICY_HIDDEN * 4;
constexpr uint32_t PACKED_BYTES = OFF_B3 + POL_N3 * 4;
static_assert(PACKED_BYTES == POLICY_PACKED_BYTES, "skyjo_policy.h out of date");
constexpr uint32_t OFF_OBS = PACKED_BYTES;                                   // 128 rows of up to POLICY_MAX_OBS bytes
constexpr uint32_t OFF_MASK = OFF_OBS + ((POLICY_TILE * POLICY_MAX_OBS + 127) / 128) * 128;
constexpr uint32_t OFF_BAR = OFF_MASK + POLICY_TILE * 26;                    // 3328 = 26 * 128
constexpr uint32_t SMEM_BYTES = OFF_BAR + 192;  // 8 + 8 mbarriers, the TMEM base address
static_assert(SMEM_BYTES <= 227 * 1024, "policy kernel shared memory");

// ---- PTX wrappers ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_load(uint32_t dst_smem, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem descriptor]; bf16 x bf16 -> fp32
__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        :
        : "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 32 lanes x 32 bits x 32 columns: thread i of the warp reads lane (warp % 4) * 32 + i, columns [col, col + 32)
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}

// K-major operand in the canonical no-swizzle layout: element (r, k) of an R x K matrix at
//   ((k / 8) * (R / 8) + r / 8) * 128 + (r % 8) * 16 + (k % 8) * 2        bytes,
// i.e. 8 x 8 core matrices of 128 contiguous bytes; consecutive 8-row groups 128 B apart (the descriptor's
// stride-dimension byte offset), the two K-halves of one K = 16 MMA (R / 8) * 128 B apart (leading-dimension
// byte offset).  cute::UMMA::SmemDescriptor: start >> 4 in [0,14), LBO >> 4 in [16,30), SBO >> 4 in [32,46),
// version 1 in [46,48), layout type 0 (no swizzle) in [61,64).
__host__ __device__ inline uint32_t kmajor_offset(uint32_t r, uint32_t k, uint32_t R) {
    return ((k >> 3) * (R >> 3) + (r >> 3)) * 128u + (r & 7u) * 16u + (k & 7u) * 2u;
}
__device__ __forceinline__ uint64_t b_desc(uint32_t smem_addr, uint32_t R) {
    const uint64_t lbo = (uint64_t)((R >> 3) * 128u) >> 4, sbo = 128u >> 4;
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | (lbo << 16) | (sbo << 32) | (1ull << 46);
}
// cute::UMMA::InstrDescriptor: c_format F32 (1) at [4,6), a/b format BF16 (1) at [7,10) / [10,13), both K-major,
// N >> 3 at [17,23), M >> 4 at [24,29)
__host__ __device__ constexpr uint32_t idesc_bf16(uint32_t M, uint32_t N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// tanh on the SFUs, everything else off them.  Measured on B200 (ncu, profiles/r2_policy_*): the XU pipe (MUFU,
// and every int <-> float / float -> bf16 conversion: I2F, F2FP) is what bounds this kernel.  MUFU.TANH occupies it
// for 16 cycles per warp instruction (half the rate of ex2 / rcp; tanh.approx.bf16x2 is issued as two of them, and
// 1 - 2 / (exp2(2 log2(e) a) + 1) through ex2 + rcp costs the same 16 cycles plus three issue slots), a conversion
// for 8.  So the conversions are done on the ALU instead: an activation is rounded to bf16 by adding 0x8000 to its
// fp32 bits and keeping the high half (PRMT packs two), an int8 observation becomes a float through the 1.5 x 2^23
// trick and is exact in its high half.  Floor: 512 tanh per env x 16 cycles / 32 lanes / 4 schedulers per SM
// = 8192 cycles per 128-env tile = 58 us per 2^18 envs at 1.965 GHz, 2.6 times the kernel's tensor-core time.
// Also measured and rejected: three of every four units as sign(a) (1 - u) / (1 + u), u = exp2(-2 log2(e) |a|)
// (one ex2: 8 XU cycles) with 1 / (1 + u) as a degree-4 minimax polynomial on the FMA pipe (2.2e-4 accurate) --
// XU time per chunk 320 instead of 512 cycles on paper, but the ten issue slots per unit of a lone warp per
// scheduler take longer than the pipe they relieve: 114 us against 89.5 (119 with two H warps per scheduler).
__device__ __forceinline__ float tanh_f32(float x) {
#ifdef SKYJO_POLICY_EXP_NOTANH  // timing experiment only: what the kernel costs without its SFU work
    return x;
#else
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#endif
}
// two fp32 -> one word of two bf16 (element 2j in the low half), round half up on the magnitude bits
__device__ __forceinline__ uint32_t pack_bf16x2_alu(float lo, float hi) {
    return __byte_perm(__float_as_uint(lo) + 0x8000u, __float_as_uint(hi) + 0x8000u, 0x7632);
}
// int8 -> fp32 without I2F: 0x4B400000 is 1.5 x 2^23, whose unit in the last place is 1
__device__ __forceinline__ float int8_to_float(int v) { return __uint_as_float(0x4B400000u + (uint32_t)v) - 12582912.0f; }

// column (relative to its 256-column region) of the bf16 pairs of hidden units [32 c, 32 c + 32): a chunk of 32
// fp32 columns compacts in place to 16 columns -- chunks 0-3 to the start of the first half of the region, chunks
// 4-7 to the start of the second half, so that each half of team H only overwrites columns it has read itself
__host__ __device__ constexpr uint32_t hidden_col(uint32_t c) {
    return (POL_HGROUPS == 1 || c < 4) ? 16u * c : 128u + 16u * (c - 4u);
}

// tanh(acc + bias) of one half (group 0: hidden units 0..127, group 1: 128..255) of the 256 fp32 columns of
// `region`, packed to bf16 pairs in place (same lanes); the next layer's MMAs over K-chunk c start as soon as all
// 128 rows have announced it on bar_h[c].
__device__ __forceinline__ void hidden_epilogue(uint32_t region, int group, const float *s_bias, float *dbg_row,
                                                uint32_t bar_h) {
    constexpr int PER = 8 / POL_HGROUPS;
#pragma unroll 1
    for (int c = PER * group; c < PER * group + PER; ++c) {
        uint32_t v[32];
        tc_ld32(region + 32 * c, v);
        tc_wait_ld();
        if (dbg_row) {
#pragma unroll
            for (int i = 0; i < 32; ++i) dbg_row[32 * c + i] = __uint_as_float(v[i]) + s_bias[32 * c + i];
        }
        uint32_t w[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float2 b = *reinterpret_cast<const float2 *>(s_bias + 32 * c + 2 * j);
            w[j] = pack_bf16x2_alu(tanh_f32(__uint_as_float(v[2 * j]) + b.x), tanh_f32(__uint_as_float(v[2 * j + 1]) + b.y));
        }
        tc_st16(region + hidden_col(c), w);
        tc_wait_st();
        tc_fence_before();
        mbar_arrive(bar_h + 8 * c);
    }
}

// Three roles, one tile (128 envs) after the other per CTA, two tiles in flight:
//   team H (4 or 8 warps)  the two hidden epilogues of every tile -- the critical path, kept busy back to back;
//   team S (4 warps)       stages the NEXT tile's observation rows into tensor memory while H is in its second epilogue,
//                          then samples the CURRENT tile's actions from the logits while H has moved on to the next tile;
//   one more warp, lane 0  issues the MMAs.
// Both teams see all 128 TMEM lanes (a warp reaches lanes 32 (warp % 4) ..).  The two 256-column halves of tensor
// memory swap roles from tile to tile (j = the CTA's tile counter, P = primary = j odd ? [256, 512) : [0, 256),
// Q = the other half):
//   x(j) as bf16 in Q[192, 240)  ->  D1(j) in P  -> h1(j) in P[0, 128) (two H groups: P[0, 64) + P[128, 192))
//   ->  D2(j) in Q  ->  h2(j) likewise in Q  ->  D3(j) (logits) in P[D3], D3 = 128 (two groups: 64)
// x(j+1) is staged into Q(j+1)[192, 240) = P(j)[192, 240) once H has compacted D1(j) (bar_e); D1(j+1) lands in
// P(j+1) = Q(j) once MMA3(j) has read h2(j) from it (the MMA warp waits for its own bar_d3); D2(j+1) overwrites
// D3(j) in Q(j+1) = P(j) once S has pulled the logits into registers (bar_c).
__global__ void __launch_bounds__(POL_THREADS, 1) policy_kernel(const __grid_constant__ PolicyParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t sm = smem_u32(smem);
    // bar_w  weights landed (tx)                       bar_in  a tile's obs / mask rows landed (tx)
    // bar_a  S -> MMA: x(j) is in tensor memory (128)   bar_d1 / bar_d2 / bar_d3  MMA -> H / H / S: a layer's MMAs completed
    // bar_h[c]  H -> MMA: chunk c of the current hidden layer is in tensor memory (128)
    // bar_e  H -> S: D1(j) has been compacted (256)     bar_c  S -> MMA: the logits of tile j are in registers (128)
    const uint32_t bar_w = sm + OFF_BAR, bar_in = bar_w + 8, bar_a = bar_w + 16, bar_d1 = bar_w + 24, bar_d2 = bar_w + 32,
                   bar_d3 = bar_w + 40, bar_e = bar_w + 48, bar_c = bar_w + 56, bar_h = bar_w + 64;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + OFF_BAR + 128);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long n_tiles = (p.B + POLICY_TILE - 1) / POLICY_TILE;
    const int D = p.D;

    if (tid == 0) {
        mbar_init(bar_w, 1);
        mbar_init(bar_in, 1);
        mbar_init(bar_a, POLICY_TILE);
        mbar_init(bar_d1, 1);
        mbar_init(bar_d2, 1);
        mbar_init(bar_d3, 1);
        mbar_init(bar_e, POL_HGROUPS * POLICY_TILE);
        mbar_init(bar_c, POLICY_TILE);
        for (int c = 0; c < 8; ++c) mbar_init(bar_h + 8 * c, POLICY_TILE);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == POL_MMA_WARP) {  // one warp allocates all of tensor memory (one CTA per SM: 210 KB of shared memory)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    // a tile whose 128 rows exist travels by bulk copy; the ragged last tile of the batch by plain loads
    auto tile_is_bulk = [&](long long tile) { return p.bulk_ok && (tile + 1) * POLICY_TILE <= p.B; };
    auto issue_tile = [&](long long tile) {
        const uint32_t ob = (uint32_t)(POLICY_TILE * D), mb = POLICY_TILE * 26u;
        mbar_expect_tx(bar_in, ob + mb);
        bulk_load(sm + OFF_OBS, p.obs + tile * POLICY_TILE * D, ob, bar_in);
        bulk_load(sm + OFF_MASK, p.mask + tile * POLICY_TILE * 26, mb, bar_in);
    };
    const long long first_tile = blockIdx.x, stride = gridDim.x;
    // timeline marks of CTA 0 (one lane per role): clock64() at every hand-over, decoded by tools/policy_trace.py
    long long *trace_role = nullptr;
    int trace_n = 0;
    auto mark = [&]() {
        if (trace_role && trace_n < POLICY_TRACE_LEN) trace_role[trace_n++] = clock64();
    };

    if (tid == 0) {  // the weight image, once per CTA (L2-resident after the first CTAs)
        mbar_expect_tx(bar_w, PACKED_BYTES);
        for (uint32_t off = 0; off < PACKED_BYTES; off += 32768u) {
            const uint32_t n = PACKED_BYTES - off < 32768u ? PACKED_BYTES - off : 32768u;
            bulk_load(sm + off, p.packed + off, n, bar_w);
        }
    }

    if (warp < POL_H_WARPS) {
        // ---- team H: the hidden epilogues ------------------------------------------------------------------------
        const int group = warp >> 2, row = tid & (POLICY_TILE - 1);
        const uint32_t lane_base = (uint32_t)(32 * (warp & 3)) << 16;
        const float *s_b1 = reinterpret_cast<const float *>(smem + OFF_B1);
        const float *s_b2 = reinterpret_cast<const float *>(smem + OFF_B2);
        if (p.trace && blockIdx.x == 0 && tid == 0) trace_role = p.trace;
        mbar_wait(bar_w, 0);
        uint32_t j = 0;
        for (long long tile = first_tile; tile < n_tiles; tile += stride, ++j) {
            const uint32_t P = tmem + lane_base + ((j & 1u) ? 256u : 0u), Q = tmem + lane_base + ((j & 1u) ? 0u : 256u);
            const long long e = tile * POLICY_TILE + row;
            const bool valid = e < p.B;
            mark();                     // H0: waiting for D1
            mbar_wait(bar_d1, j & 1u);  // D1 = x W1^T
            tc_fence_after();
            mark();                     // H1: first epilogue starts
            hidden_epilogue(P, group, s_b1, (p.dbg1 && valid) ? p.dbg1 + e * POLICY_HIDDEN : nullptr, bar_h);
            mbar_arrive(bar_e);         // (ordered behind the tcgen05 fence of the last chunk)
            mark();                     // H2: first epilogue done, waiting for D2
            mbar_wait(bar_d2, j & 1u);  // D2 = h1 W2^T
            tc_fence_after();
            mark();                     // H3: second epilogue starts
            hidden_epilogue(Q, group, s_b2, (p.dbg2 && valid) ? p.dbg2 + e * POLICY_HIDDEN : nullptr, bar_h);
            mark();                     // H4: second epilogue done
        }
    } else if (warp < POL_MMA_WARP) {
        // ---- team S: operand staging one tile ahead, output epilogue one tile behind -------------------------------
        const int row = tid - POL_HGROUPS * POLICY_TILE;
        const uint32_t lane_base = (uint32_t)(32 * (warp & 3)) << 16;
        const float *s_b3 = reinterpret_cast<const float *>(smem + OFF_B3);
        const int8_t *s_obs = reinterpret_cast<const int8_t *>(smem + OFF_OBS);
        const int8_t *s_mask = reinterpret_cast<const int8_t *>(smem + OFF_MASK);
        uint32_t ph_in = 0;
        if (row == 0 && first_tile < n_tiles && tile_is_bulk(first_tile)) issue_tile(first_tile);

        // x(j) of `tile` -> Q(j)[192, 240) as bf16 pairs (exact: -24 .. 127; zero beyond D); returns the row's legal bits
        auto stage = [&](long long tile, uint32_t j) -> uint32_t {
            const long long e = tile * POLICY_TILE + row;
            const bool valid = e < p.B;
            if (tile_is_bulk(tile)) {
                mbar_wait(bar_in, ph_in);
                ph_in ^= 1u;
            } else {
                const long long rows = p.B - tile * POLICY_TILE < POLICY_TILE ? p.B - tile * POLICY_TILE : POLICY_TILE;
                for (int i = row; i < rows * D; i += POLICY_TILE) smem[OFF_OBS + i] = (uint8_t)p.obs[tile * POLICY_TILE * D + i];
                for (int i = row; i < rows * 26; i += POLICY_TILE) smem[OFF_MASK + i] = (uint8_t)p.mask[tile * POLICY_TILE * 26 + i];
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
            const uint32_t Qx = tmem + lane_base + ((j & 1u) ? 0u : 256u) + 192u;
            const int8_t *orow = s_obs + row * D;
#pragma unroll
            for (int g = 0; g < POLICY_K1 / 32; ++g) {
                uint32_t w[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const int k = 32 * g + 2 * q;
                    const float x0 = (valid && k < D) ? int8_to_float(orow[k]) : 0.f;
                    const float x1 = (valid && k + 1 < D) ? int8_to_float(orow[k + 1]) : 0.f;
                    w[q] = __byte_perm(__float_as_uint(x0), __float_as_uint(x1), 0x7632);  // exact in bf16
                }
                tc_st16(Qx + 16 * g, w);
            }
            uint32_t legal = 0;
            if (valid) {
                const int8_t *mr = s_mask + row * 26;
#pragma unroll
                for (int a = 0; a < 26; ++a) legal |= (mr[a] != 0 ? 1u : 0u) << a;
            }
            tc_wait_st();
            tc_fence_before();
            mbar_arrive(bar_a);
            // every row of the tile has been read: the next tile's rows may land
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (row == 0 && tile + stride < n_tiles && tile_is_bulk(tile + stride)) issue_tile(tile + stride);
            return legal;
        };

        if (p.trace && blockIdx.x == 0 && row == 0) trace_role = p.trace + POLICY_TRACE_LEN;
        mbar_wait(bar_w, 0);  // b3
        uint32_t legal = 0, j = 0;
        if (first_tile < n_tiles) legal = stage(first_tile, 0);
        for (long long tile = first_tile; tile < n_tiles; tile += stride, ++j) {
            uint32_t legal_next = 0;
            mark();                        // S0: waiting for the staging slot
            if (tile + stride < n_tiles) {
                mbar_wait(bar_e, j & 1u);  // D1(j) compacted: P(j)[192, 240) = Q(j+1)[192, 240) is free
                tc_fence_after();
                mark();                    // S1: staging x(j+1)
                legal_next = stage(tile + stride, j + 1);
            } else {
                mark();
            }
            const long long e = tile * POLICY_TILE + row;
            const bool valid = e < p.B;
            mark();                        // S2: staged, waiting for the logits
            mbar_wait(bar_d3, j & 1u);  // D3 = h2 W3^T
            tc_fence_after();
            mark();                        // S3: sampling
            uint32_t v[32];
            tc_ld32(tmem + lane_base + ((j & 1u) ? 256u : 0u) + POL_D3_COL, v);
            tc_wait_ld();
            tc_fence_before();
            mbar_arrive(bar_c);
            if (valid) {
                if (p.value) {
                    p.value[e] = __uint_as_float(v[0]) + s_b3[0];
                } else {
                    float l[26];
#pragma unroll
                    for (int a = 0; a < 26; ++a) l[a] = __uint_as_float(v[a]) + s_b3[a];
                    if (p.logits) {
#pragma unroll
                        for (int a = 0; a < 26; ++a) p.logits[e * 26 + a] = l[a];
                    }
                    if (p.actions) {
                        if (legal == 0) {  // cannot happen for a live env; keep the step well defined
                            p.actions[e] = 255;
                            if (p.logp) p.logp[e] = 0.f;
                            if (p.entropy) p.entropy[e] = 0.f;
                        } else {
                            int act;
                            float lp, ent;
                            sample_masked(l, legal, p.seed, p.first_env + (unsigned long long)e, p.t, act, lp, ent);
                            p.actions[e] = (uint8_t)act;
                            if (p.logp) p.logp[e] = lp;
                            if (p.entropy) p.entropy[e] = ent;
                        }
                    }
                }
            }
            legal = legal_next;
            mark();                        // S4: sampled
        }
    } else {
        // ---- the MMA warp: 6 + 16 + 16 tcgen05.mma per tile, all issued by lane 0; A from tensor memory, B from shared
        // memory.  The other lanes wait at the __syncwarp below.  Chunks are taken in the order the two halves of team
        // H produce them (0, 4, 1, 5, ...), so that the MMAs of a layer end right behind its epilogue.
        constexpr uint32_t I256 = idesc_bf16(POLICY_TILE, POLICY_HIDDEN), I32 = idesc_bf16(POLICY_TILE, POL_N3);
        constexpr uint32_t KSTEP_H = 2u * (POLICY_HIDDEN / 8) * 128u;  // bytes between K = 16 slices of W1 / W2
        constexpr uint32_t KSTEP_3 = 2u * (POL_N3 / 8) * 128u;
        if (lane == 0) {
            uint32_t ph_h = 0, j = 0;
            const uint32_t k1_steps = (uint32_t)(D + 15) / 16u;  // 5 of 6 for the 67-byte rows of four players
            if (p.trace && blockIdx.x == 0) trace_role = p.trace + 2 * POLICY_TRACE_LEN;
            mbar_wait(bar_w, 0);
            for (long long tile = first_tile; tile < n_tiles; tile += stride, ++j) {
                const uint32_t P = tmem + ((j & 1u) ? 256u : 0u), Q = tmem + ((j & 1u) ? 0u : 256u);
                mark();                                        // M0: waiting for x(j)
                // No wait for MMA3(j-1), which reads h2(j-1) out of P(j): tcgen05.mma executes in issue order, and
                // the raw D2(j-1) columns were consumed by team H before it announced the chunks MMA3(j-1) ran on.
                mbar_wait(bar_a, j & 1u);
                tc_fence_after();
                mark();                                        // M1: issuing MMA1
#pragma unroll
                for (uint32_t k = 0; k < POLICY_K1 / 16; ++k)  // K-steps beyond the row length multiply zeros: skipped
                    if (k < k1_steps) tc_mma_ts(P, Q + 192u + 8u * k, b_desc(sm + OFF_W1 + k * KSTEP_H, POLICY_HIDDEN), I256, k);
                tc_commit(bar_d1);
                // layers 2 and 3 run behind the epilogue that produces their operand: two MMAs (K = 32) per announced chunk
#pragma unroll 1
                for (uint32_t i = 0; i < 8; ++i) {
                    const uint32_t c = POL_HGROUPS == 1 ? i : (i >> 1) + 4u * (i & 1u);
                    mbar_wait(bar_h + 8 * c, ph_h);
                    if (i == 0 && j > 0) mbar_wait(bar_c, (j - 1u) & 1u);  // the logits of tile j-1 have left Q(j)[64, 96)
                    tc_fence_after();
#pragma unroll
                    for (uint32_t k = 2 * c; k < 2 * c + 2; ++k)
                        tc_mma_ts(Q, P + hidden_col(c) + 8u * (k & 1u), b_desc(sm + OFF_W2 + k * KSTEP_H, POLICY_HIDDEN), I256, i + (k & 1u));
                    if (i == 7) tc_commit(bar_d2);
                    mark();                                    // M2..M9: MMA2 chunk issued
                }
                ph_h ^= 1u;
#pragma unroll 1
                for (uint32_t i = 0; i < 8; ++i) {
                    const uint32_t c = POL_HGROUPS == 1 ? i : (i >> 1) + 4u * (i & 1u);
                    mbar_wait(bar_h + 8 * c, ph_h);
                    tc_fence_after();
#pragma unroll
                    for (uint32_t k = 2 * c; k < 2 * c + 2; ++k)
                        tc_mma_ts(P + POL_D3_COL, Q + hidden_col(c) + 8u * (k & 1u), b_desc(sm + OFF_W3 + k * KSTEP_3, POL_N3), I32, i + (k & 1u));
                    if (i == 7) tc_commit(bar_d3);
                    mark();                                    // M10..M17: MMA3 chunk issued
                }
                ph_h ^= 1u;
            }
        }
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == POL_MMA_WARP) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

// fp32 torch.nn.Linear weights (W[out][in], row-major) -> the kernel's shared-memory image
__global__ void policy_pack_kernel(const float *w1, const float *b1, const float *w2, const float *b2, const float *w3,
                                   const float *b3, int D, int n_out, uint8_t *packed) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    __nv_bfloat16 *W1 = reinterpret_cast<__nv_bfloat16 *>(packed + OFF_W1);
    __nv_bfloat16 *W2 = reinterpret_cast<__nv_bfloat16 *>(packed + OFF_W2);
    __nv_bfloat16 *W3 = reinterpret_cast<__nv_bfloat16 *>(packed + OFF_W3);
    if (i < POLICY_K1 * POLICY_HIDDEN) {
        const int n = i / POLICY_K1, k = i % POLICY_K1;
        W1[kmajor_offset(n, k, POLICY_HIDDEN) / 2] = __float2bfloat16_rn(k < D ? w1[n * D + k] : 0.f);
    }
    if (i < POLICY_HIDDEN * POLICY_HIDDEN) {
        const int n = i / POLICY_HIDDEN, k = i % POLICY_HIDDEN;
        W2[kmajor_offset(n, k, POLICY_HIDDEN) / 2] = __float2bfloat16_rn(w2[n * POLICY_HIDDEN + k]);
    }
    if (i < POL_N3 * POLICY_HIDDEN) {
        const int n = i / POLICY_HIDDEN, k = i % POLICY_HIDDEN;
        W3[kmajor_offset(n, k, POL_N3) / 2] = __float2bfloat16_rn(n < n_out ? w3[n * POLICY_HIDDEN + k] : 0.f);
    }
    if (i < POLICY_HIDDEN) {
        reinterpret_cast<float *>(packed + OFF_B1)[i] = b1[i];
        reinterpret_cast<float *>(packed + OFF_B2)[i] = b2[i];
    }
    if (i < POL_N3) reinterpret_cast<float *>(packed + OFF_B3)[i] = i < n_out ? b3[i] : 0.f;
}

cudaError_t launch_policy_pack(const float *w1, const float *b1, const float *w2, const float *b2, const float *w3,
                               const float *b3, int D, int n_out, void *packed, cudaStream_t s) {
    const int n = POLICY_HIDDEN * POLICY_HIDDEN;
    policy_pack_kernel<<<(n + 255) / 256, 256, 0, s>>>(w1, b1, w2, b2, w3, b3, D, n_out, (uint8_t *)packed);
    return cudaGetLastError();
}

cudaError_t launch_policy(const PolicyParams &p, int sm_count, cudaStream_t s) {
    const cudaError_t e = cudaFuncSetAttribute(policy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    if (e != cudaSuccess) return e;
    const long long n_tiles = (p.B + POLICY_TILE - 1) / POLICY_TILE;
    const unsigned grid = (unsigned)(n_tiles < sm_count ? n_tiles : sm_count);
    policy_kernel<<<grid, POL_THREADS, SMEM_BYTES, s>>>(p);
    return cudaGetLastError();
}

}  // namespace skyjo
