// skyjo_hostio.cuh -- the wire format of skyjo_step_host (host buffers in, host buffers out).
//
// The end-to-end rate of the host entry is set by the device-to-host link (PCIe Gen5 x16,
// ~54 GB/s measured), so the bytes that cross it are cut to what carries information:
//   * action mask (26 B) + agent (1 B) + done (1 B) travel as ONE 32-bit word per env:
//       bits 0..25 legal-action bits (bit a = action a, _jit_action_mask skyjo.py:201-224),
//       bits 26..27 done code, bits 28..31 agent (expected_action[0]);
//     the host expands it into the caller's int8 buffers while the observation copy is still
//     in flight (a few worker threads; the expansion is pure byte spreading).
//   * rewards are zero except in the step that ends an episode (skyjo_env.py:242-247): the pack
//     kernel compacts the rows of the envs whose done code is non-zero into a host-mapped
//     pinned buffer ({env index, N doubles} per entry) and the host scatters them; rows written
//     by the previous call are re-zeroed first.  If more envs finish than the buffer holds
//     (mass illegal actions) the call falls back to copying the dense reward tensor.
//   * observations (D bytes per env) are copied as they are, straight into the caller's buffer (wire mode 0).
//     Wire mode 1 sends COMPACT RECORDS instead (mode 2, the default where the CPU has AVX-512, for a measured share of
//     the env ranges -- skyjo_capi.cu skyjo_step_host): every byte of an observation row (skyjo.py:180-190) is
//     one of a few small symbols -- a card is -2..12, 15 (hidden) or -14 (removed column), a histogram bin is a
//     count <= 15 (the value-0 bin, which receives three zeros per column removal, up to a byte), the discard
//     top is -3..12 -- so a row of D = 19 + 12 R bytes packs into 12 + 6 R + ceil(R / 2) bytes
//     (obs_record_bytes): two symbols per byte, one removed-column flag nibble per card row.  compact_obs_kernel
//     packs the published rows on the device, the host expands them into the caller's int8[B, D] buffer with
//     16-byte shuffles (pshufb as the nibble -> card value table).  A row that holds anything else (it cannot,
//     by the state layout: 4-bit bins) raises a flag and the call falls back to the dense copy.
//     (AVX-512 hosts: byte gathers and streaming stores, skyjo_hostsimd.cpp.)  Measured (B200 boxes, N = 4, 2^20
//     envs): 42 instead of 71 B per env on the link, but the expansion writes 95 B per env through the CPUs where the
//     copy engine wrote them for free, and the two share the host's memory system: all ranges compact 5.8-7.0e8
//     env-steps/s, all raw 7.2e8, four of eight compact 7.7-8.0e8 at 1 GPU; at 8 GPUs raw wins.
// The batch is processed in up to HOSTIO_MAX_CHUNKS env ranges (step kernel, pack kernel and copies
// per range), so that the link is already busy with range c while the GPU steps range c + 1.
// N = 4, direct observations: 67 + 4 = 71 B per env-step (38 + 4 = 42 B in wire mode 1) instead of 127 B.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#if defined(__linux__)
#include <pthread.h>
#include <sched.h>
#endif
#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "skyjo_state.cuh"

namespace skyjo {

constexpr int PACK_DONE_SH = 26, PACK_AGENT_SH = 28;
constexpr int HOSTIO_MAX_CHUNKS = 8;

// One thread per env: the packed word from the state planes (the same legal_bits the fused step
// kernel encoded the mask row from) and the compaction of finished envs' reward rows.
// `entries` is host-mapped pinned memory: entry i = (1 + N) doubles, the first holding the env
// index as uint64.
__global__ void __launch_bounds__(256) pack_host_kernel(const U128 *planes, long long Bpad, long long e_begin,
                                                        long long B, int N,
                                                        const uint8_t *done, const double *reward,
                                                        uint32_t *packed, unsigned int *counter,
                                                        double *entries, unsigned int cap) {
    const long long e = e_begin + (long long)blockIdx.x * 256 + threadIdx.x;
    if (e >= B) return;  // B = end of this launch's env range
    const U128 P0 = ld128(planes + e);
    const uint64_t hdr = pack64(P0.x, P0.y);
    const uint32_t cur = (uint32_t)(hdr >> HDR_CUR_SH) & 0xFu;
    const U128 T = ld128(planes + (long long)(1 + cur) * Bpad + e);
    Row r{T.x, T.y, T.z, T.w};
    const uint32_t lb = legal_bits(row_hidden(r), row_flags(r), (hdr & HDR_PHASE) != 0);
    const uint32_t d = done[e];
    packed[e] = lb | (d << PACK_DONE_SH) | (cur << PACK_AGENT_SH);
    if (d && entries) {
        const unsigned int i = atomicAdd(counter, 1u);
        if (i < cap) {
            double *dst = entries + (size_t)i * (size_t)(1 + N);
            reinterpret_cast<unsigned long long *>(dst)[0] = (unsigned long long)e;
            for (int q = 0; q < N; ++q) dst[1 + q] = reward[e * N + q];
        }
    }
}


// ---- small batches: one block publishes everything into host-mapped memory ---------------------------
// For a handful of envs (the single-env PettingZoo view, skyjo_rl_b200/aec.py) a call is latency, not bandwidth:
// the step kernel reads the actions through a host-mapped pointer and this one-block kernel writes the
// published outputs straight into host-mapped pinned memory, followed by a sequence number the host polls --
// no copy-engine transfers, no event, and the refill deal is launched only when an env finished.
// Staging layout (bytes): [0, 16) seq + finished count, then actions B (padded to 16), obs B x D, mask B x 26,
// agent B, done B (each padded to 16), reward B x N doubles.
constexpr long long SMALL_HOST_MAX = 256;
__host__ __device__ inline size_t small_pad16(size_t n) { return (n + 15) & ~(size_t)15; }
struct SmallLayout {
    size_t act, obs, mask, agent, done, reward, total;
};
__host__ __device__ inline SmallLayout small_layout(size_t B, size_t D, size_t N) {
    SmallLayout L;
    L.act = 16;
    L.obs = L.act + small_pad16(B);
    L.mask = L.obs + small_pad16(B * D);
    L.agent = L.mask + small_pad16(B * 26);
    L.done = L.agent + small_pad16(B);
    L.reward = L.done + small_pad16(B);
    L.total = L.reward + B * N * 8;
    return L;
}
__global__ void __launch_bounds__(256) publish_small_kernel(const int8_t *obs, const int8_t *mask, const int8_t *agent,
                                                            const uint8_t *done, const double *reward, long long B,
                                                            int D, int N, uint8_t *stage, unsigned int seq) {
    const SmallLayout L = small_layout((size_t)B, (size_t)D, (size_t)N);
    const int t = threadIdx.x;
    __shared__ unsigned int s_fin;
    if (t == 0) s_fin = 0;
    __syncthreads();
    for (long long i = t; i < B * D; i += 256) stage[L.obs + i] = (uint8_t)obs[i];
    for (long long i = t; i < B * 26; i += 256) stage[L.mask + i] = (uint8_t)mask[i];
    unsigned int fin = 0;
    for (long long i = t; i < B; i += 256) {
        stage[L.agent + i] = (uint8_t)agent[i];
        stage[L.done + i] = done[i];
        fin += done[i] != 0;
    }
    double *rw = reinterpret_cast<double *>(stage + L.reward);
    for (long long i = t; i < B * N; i += 256) rw[i] = reward[i];
    if (fin) atomicAdd(&s_fin, fin);
    __threadfence_system();
    __syncthreads();
    if (t == 0) {
        reinterpret_cast<volatile unsigned int *>(stage)[1] = s_fin;
        __threadfence_system();
        reinterpret_cast<volatile unsigned int *>(stage)[0] = seq;  // last: the host polls this word
    }
}

// ---- compact observation records ------------------------------------------------------------------
// Row layout (skyjo.py:180-190): [0] min open sum, [1] min hidden count, [2..16] 15-bin histogram (bin k = value
// k - 2), [17] discard top, [18] hand card, [19 + 12 r + i] slot i of card row r (R = N rows, or the own row).
// Record:  [0] row[0]   [1] row[1] | (top + 3) << 4   [2] hand code   [3] row[4] (the value-0 bin, raw byte)
//          [4..11] histogram nibbles (nibble k = bin k, k = 2 and k = 15 unused = 0)
//          [12 .. 12 + 6R) card codes, two per byte: value + 2 for -2..12, 15 = hidden, 0 where the column is removed
//          [12 + 6R ..) removed-column flags, one nibble per card row (bit c = column c shows -14)
__host__ __device__ inline int obs_record_bytes(int D) {
    const int R = (D - 19) / 12;
    return 12 + 6 * R + (R + 1) / 2;
}

// code of one card slot; 16 = the removed marker (-14), 255 = not encodable
__host__ __device__ inline uint32_t card_code(int v) {
    if (v == 15) return 15u;
    if (v == -14) return 16u;
    return (v >= -2 && v <= 12) ? (uint32_t)(v + 2) : 255u;
}

// Packs one observation row; false if the row holds a byte outside the symbol sets above.
__host__ __device__ inline bool pack_obs_record(const int8_t *row, int D, uint8_t *rec) {
    const int R = (D - 19) / 12;
    bool ok = true;
    const int nh = row[1], top = row[17] + 3, hand = row[18];
    const uint32_t hc = hand == 15 ? 15u : (uint32_t)(hand + 2);
    ok = ok && nh >= 0 && nh <= 15 && top >= 0 && top <= 15 && (hand == 15 || (hand >= -2 && hand <= 12));
    rec[0] = (uint8_t)row[0];
    rec[1] = (uint8_t)((nh & 15) | ((top & 15) << 4));
    rec[2] = (uint8_t)(hc & 15u);
    rec[3] = (uint8_t)row[4];
    for (int j = 0; j < 8; ++j) {
        const int k0 = 2 * j, k1 = 2 * j + 1;
        const int b0 = (k0 == 2) ? 0 : row[2 + k0];
        const int b1 = (k1 == 15) ? 0 : row[2 + k1];
        ok = ok && b0 >= 0 && b0 <= 15 && b1 >= 0 && b1 <= 15;
        rec[4 + j] = (uint8_t)((b0 & 15) | ((b1 & 15) << 4));
    }
    uint8_t *cards = rec + 12, *flags = rec + 12 + 6 * R;
    for (int r = 0; r < R; ++r) {
        const int8_t *c = row + 19 + 12 * r;
        uint32_t fl = 0;
        for (int col = 0; col < 4; ++col) {
            const uint32_t a = card_code(c[3 * col]), b = card_code(c[3 * col + 1]), d = card_code(c[3 * col + 2]);
            const uint32_t removed = (a == 16u) + (b == 16u) + (d == 16u);
            ok = ok && (removed == 0u || removed == 3u) && a != 255u && b != 255u && d != 255u;
            fl |= (removed == 3u ? 1u : 0u) << col;
        }
        for (int j = 0; j < 6; ++j) {
            const uint32_t a = card_code(c[2 * j]) & 15u, b = card_code(c[2 * j + 1]) & 15u;  // 16 -> 0
            cards[6 * r + j] = (uint8_t)(a | (b << 4));
        }
        if (r & 1)
            flags[r >> 1] |= (uint8_t)(fl << 4);
        else
            flags[r >> 1] = (uint8_t)fl;
    }
    return ok;
}

// One thread per env of [e_begin, e_end): packs the published observation row into its record.
__global__ void __launch_bounds__(256) compact_obs_kernel(const int8_t *obs, long long e_begin, long long e_end, int D,
                                                          int RB, uint8_t *rec, unsigned int *bad) {
    const long long e = e_begin + (long long)blockIdx.x * 256 + threadIdx.x;
    if (e >= e_end) return;
    if (!pack_obs_record(obs + e * D, D, rec + e * RB)) atomicOr(bad, 1u);
}

// ---- host side --------------------------------------------------------------------------------
// 8 bits -> 8 bytes of 0/1 (byte i = bit i)
static inline uint64_t spread8(uint32_t b) {
    const uint64_t x = ((uint64_t)(b & 0xFFu) * 0x0101010101010101ull) & 0x8040201008040201ull;
    return ((x + 0x7F7F7F7F7F7F7F7Full) >> 7) & 0x0101010101010101ull;
}

static inline void expand_packed(const uint32_t *packed, long long e0, long long e1, int8_t *mask, int8_t *agent,
                                 uint8_t *done) {
    for (long long e = e0; e < e1; ++e) {
        const uint32_t p = packed[e];
        if (mask) {
            const uint64_t m0 = spread8(p), m1 = spread8(p >> 8), m2 = spread8(p >> 16);
            const uint16_t m3 = (uint16_t)(spread8(p >> 24) & 0xFFFFu);
            int8_t *row = mask + e * 26;
            memcpy(row, &m0, 8);
            memcpy(row + 8, &m1, 8);
            memcpy(row + 16, &m2, 8);
            memcpy(row + 24, &m3, 2);
        }
        if (agent) agent[e] = (int8_t)(p >> PACK_AGENT_SH);
        if (done) done[e] = (uint8_t)((p >> PACK_DONE_SH) & 3u);
    }
}


// Expands records [e0, e1) into obs rows (the inverse of pack_obs_record), portable version.
static inline void expand_obs_scalar(const uint8_t *rec, long long e0, long long e1, int D, int8_t *obs) {
    const int R = (D - 19) / 12, RB = obs_record_bytes(D);
    for (long long e = e0; e < e1; ++e) {
        const uint8_t *r = rec + e * RB;
        int8_t *o = obs + e * D;
        for (int j = 0; j < 8; ++j) {
            o[2 + 2 * j] = (int8_t)(r[4 + j] & 15);
            if (2 * j + 1 < 15) o[3 + 2 * j] = (int8_t)(r[4 + j] >> 4);
        }
        o[0] = (int8_t)r[0];
        o[1] = (int8_t)(r[1] & 15);
        o[4] = (int8_t)r[3];
        o[17] = (int8_t)((r[1] >> 4) - 3);
        o[18] = (int8_t)((r[2] & 15) == 15 ? 15 : (r[2] & 15) - 2);
        const uint8_t *cards = r + 12, *flags = r + 12 + 6 * R;
        for (int i = 0; i < 6 * R; ++i) {
            const int a = cards[i] & 15, b = cards[i] >> 4;
            o[19 + 2 * i] = (int8_t)(a == 15 ? 15 : a - 2);
            o[20 + 2 * i] = (int8_t)(b == 15 ? 15 : b - 2);
        }
        for (int q = 0; q < R; ++q) {
            const uint32_t fl = (flags[q >> 1] >> (4 * (q & 1))) & 15u;
            for (int col = 0; fl && col < 4; ++col)
                if (fl >> col & 1u) o[19 + 12 * q + 3 * col] = o[20 + 12 * q + 3 * col] = o[21 + 12 * q + 3 * col] = -14;
        }
    }
}

#if defined(__x86_64__)
// 8 record bytes (16 nibbles, low nibble first) -> 16 bytes, through a 16-entry table
__attribute__((target("ssse3"))) static inline __m128i nibbles16(const uint8_t *src, __m128i lut) {
    const __m128i x = _mm_loadl_epi64(reinterpret_cast<const __m128i *>(src));
    const __m128i m = _mm_set1_epi8(0x0F);
    const __m128i idx = _mm_unpacklo_epi8(_mm_and_si128(x, m), _mm_and_si128(_mm_srli_epi16(x, 4), m));
    return _mm_shuffle_epi8(lut, idx);
}

__attribute__((target("ssse3"))) static void expand_obs_ssse3(const uint8_t *rec, long long e0, long long e1, int D,
                                                              int8_t *obs) {
    const int R = (D - 19) / 12, RB = obs_record_bytes(D), NC = 6 * R;
    const __m128i ident = _mm_setr_epi8(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15);
    const __m128i cardv = _mm_setr_epi8(-2, -1, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 15);
    for (long long e = e0; e < e1; ++e) {
        const uint8_t *r = rec + e * RB;
        int8_t *o = obs + e * D;
        // histogram: 16 nibbles -> o[2..17]; o[4] (value-0 bin) and o[17] (top) are rewritten below
        _mm_storeu_si128(reinterpret_cast<__m128i *>(o + 2), nibbles16(r + 4, ident));
        o[0] = (int8_t)r[0];
        o[1] = (int8_t)(r[1] & 15);
        o[4] = (int8_t)r[3];
        o[17] = (int8_t)((r[1] >> 4) - 3);
        o[18] = (int8_t)((r[2] & 15) == 15 ? 15 : (r[2] & 15) - 2);
        const uint8_t *cards = r + 12;
        int i = 0;
        for (; i + 8 <= NC; i += 8)
            _mm_storeu_si128(reinterpret_cast<__m128i *>(o + 19 + 2 * i), nibbles16(cards + i, cardv));
        if (i < NC) {  // 6 R is not a multiple of 8: the last 2 / 4 / 6 record bytes
            uint64_t tail = 0;
            memcpy(&tail, cards + i, (size_t)(NC - i));
            alignas(16) int8_t tmp[16];
            _mm_store_si128(reinterpret_cast<__m128i *>(tmp), nibbles16(reinterpret_cast<const uint8_t *>(&tail), cardv));
            memcpy(o + 19 + 2 * i, tmp, (size_t)(2 * (NC - i)));
        }
        const uint8_t *flags = r + 12 + NC;
        for (int q = 0; q < R; q += 2) {
            uint32_t fl = flags[q >> 1];
            if (!fl) continue;
            for (int h = 0; h < 2; ++h, fl >>= 4)
                for (int col = 0; col < 4; ++col)
                    if (fl >> col & 1u) {
                        int8_t *c = o + 19 + 12 * (q + h) + 3 * col;
                        c[0] = c[1] = c[2] = -14;
                    }
        }
    }
}
#endif

static inline void expand_obs_records(const uint8_t *rec, long long e0, long long e1, int D, int8_t *obs) {
#if defined(__x86_64__)
    static const bool have_ssse3 = __builtin_cpu_supports("ssse3");
    if (have_ssse3) return expand_obs_ssse3(rec, e0, e1, D, obs);
#endif
    expand_obs_scalar(rec, e0, e1, D, obs);
}

// Fork-join pool for the host-side expansion (created on the first skyjo_step_host call).  The workers are
// persistent and, unless SKYJO_HOST_PIN=0, pinned one per core to this rank's slice of the CPUs the process may
// run on (slice LOCAL_RANK of LOCAL_WORLD_SIZE under torchrun), so the ranks of a box do not migrate over each
// other's cores.  A worker spins briefly on the generation counter before it sleeps: a call runs the pool once
// per env range, a few hundred microseconds apart, and a futex wake-up costs tens of microseconds each time.
class HostPool {
  public:
    explicit HostPool(int n) : n_(n < 1 ? 1 : n) {
        if (const char *g = getenv("SKYJO_HOST_SPIN")) spin_ = atoi(g) < 0 ? 0 : atoi(g);  // experiment knob
        std::vector<int> cpus = rank_cpus();
        for (int i = 1; i < n_; ++i) {
            workers_.emplace_back([this, i] { loop(i); });
#if defined(__linux__)
            if (!cpus.empty()) {
                cpu_set_t set;
                CPU_ZERO(&set);
                CPU_SET(cpus[(size_t)i % cpus.size()], &set);
                pthread_setaffinity_np(workers_.back().native_handle(), sizeof(set), &set);
            }
#endif
        }
    }
    ~HostPool() {
        {
            std::lock_guard<std::mutex> g(m_);
            stop_ = true;
            gen_.fetch_add(1, std::memory_order_release);
        }
        cv_.notify_all();
        for (auto &t : workers_) t.join();
    }
    int size() const { return n_; }
    // the CPUs of this rank's slice (empty = do not pin)
    static std::vector<int> rank_cpus() {
        std::vector<int> out;
#if defined(__linux__)
        if (const char *g = getenv("SKYJO_HOST_PIN"))
            if (atoi(g) == 0) return out;
        cpu_set_t set;
        CPU_ZERO(&set);
        if (sched_getaffinity(0, sizeof(set), &set) != 0) return out;
        std::vector<int> allowed;
        for (int c = 0; c < CPU_SETSIZE; ++c)
            if (CPU_ISSET(c, &set)) allowed.push_back(c);
        int ranks = 1, rank = 0;
        if (const char *g = getenv("LOCAL_WORLD_SIZE")) ranks = atoi(g) > 0 ? atoi(g) : 1;
        if (const char *g = getenv("LOCAL_RANK")) rank = atoi(g) >= 0 ? atoi(g) : 0;
        if (rank >= ranks) rank = 0;
        const size_t m = allowed.size(), lo = m * (size_t)rank / (size_t)ranks, hi = m * (size_t)(rank + 1) / (size_t)ranks;
        for (size_t k = lo; k < hi; ++k) out.push_back(allowed[k]);
#endif
        return out;
    }
    // runs fn(part, parts) for part = 0..n-1, part 0 on the calling thread; returns when all are done
    void run(const std::function<void(int, int)> &fn) {
        if (n_ == 1) {
            fn(0, 1);
            return;
        }
        {
            std::lock_guard<std::mutex> g(m_);
            fn_ = &fn;
            pending_.store(n_ - 1, std::memory_order_relaxed);
            gen_.fetch_add(1, std::memory_order_release);
        }
        cv_.notify_all();
        fn(0, n_);
        for (int spin = 0; spin < 3 * spin_ && pending_.load(std::memory_order_acquire) != 0; ++spin) cpu_relax();
        if (pending_.load(std::memory_order_acquire) != 0) {
            std::unique_lock<std::mutex> g(m_);
            done_cv_.wait(g, [this] { return pending_.load(std::memory_order_acquire) == 0; });
        }
        fn_ = nullptr;
    }

  private:
    static void cpu_relax() {
#if defined(__x86_64__)
        _mm_pause();
#endif
    }
    void loop(int i) {
        unsigned long long seen = 0;
        for (;;) {
            for (int spin = 0; spin < spin_ && gen_.load(std::memory_order_acquire) == seen; ++spin) cpu_relax();
            const std::function<void(int, int)> *fn;
            {
                std::unique_lock<std::mutex> g(m_);
                cv_.wait(g, [&] { return gen_.load(std::memory_order_acquire) != seen; });
                seen = gen_.load(std::memory_order_acquire);
                if (stop_) return;
                fn = fn_;
            }
            (*fn)(i, n_);
            if (pending_.fetch_sub(1, std::memory_order_acq_rel) == 1) {
                std::lock_guard<std::mutex> g(m_);
                done_cv_.notify_one();
            }
        }
    }
    int n_;
    int spin_ = 1500;  // pause iterations (~100 us) a worker polls before it sleeps
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_, done_cv_;
    const std::function<void(int, int)> *fn_ = nullptr;
    std::atomic<unsigned long long> gen_{0};
    std::atomic<int> pending_{0};
    bool stop_ = false;
};

}  // namespace skyjo
