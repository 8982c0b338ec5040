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
//   * observations (D bytes per env) are copied as they are.
// The batch is processed in up to HOSTIO_MAX_CHUNKS env ranges (step kernel, pack kernel and copies
// per range), so that the link is already busy with range c while the GPU steps range c + 1.
// N = 4, direct observations: 67 + 4 = 71 B per env-step instead of 127 B.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

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

// Minimal fork-join pool for the host-side expansion (created on the first skyjo_step_host call).
class HostPool {
  public:
    explicit HostPool(int n) : n_(n < 1 ? 1 : n) {
        for (int i = 1; i < n_; ++i) workers_.emplace_back([this, i] { loop(i); });
    }
    ~HostPool() {
        {
            std::lock_guard<std::mutex> g(m_);
            stop_ = true;
            ++gen_;
        }
        cv_.notify_all();
        for (auto &t : workers_) t.join();
    }
    int size() const { return n_; }
    // runs fn(part, parts) for part = 0..n-1, part 0 on the calling thread; returns when all are done
    void run(const std::function<void(int, int)> &fn) {
        if (n_ == 1) {
            fn(0, 1);
            return;
        }
        {
            std::lock_guard<std::mutex> g(m_);
            fn_ = &fn;
            pending_ = n_ - 1;
            ++gen_;
        }
        cv_.notify_all();
        fn(0, n_);
        std::unique_lock<std::mutex> g(m_);
        done_cv_.wait(g, [this] { return pending_ == 0; });
        fn_ = nullptr;
    }

  private:
    void loop(int i) {
        unsigned long long seen = 0;
        for (;;) {
            const std::function<void(int, int)> *fn;
            {
                std::unique_lock<std::mutex> g(m_);
                cv_.wait(g, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
                fn = fn_;
            }
            (*fn)(i, n_);
            {
                std::lock_guard<std::mutex> g(m_);
                if (--pending_ == 0) done_cv_.notify_one();
            }
        }
    }
    int n_;
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_, done_cv_;
    const std::function<void(int, int)> *fn_ = nullptr;
    unsigned long long gen_ = 0;
    int pending_ = 0;
    bool stop_ = false;
};

}  // namespace skyjo
