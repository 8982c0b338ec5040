// skyjo_hostsimd.cpp -- AVX-512 / non-temporal-store versions of the host half of skyjo_step_host's wire format
// (csrc/skyjo_hostio.cuh has the format and the portable code).  Host compiler only.
//
// Why: the host side of the entry fills 127 B per env-step of caller memory.  Plain stores read every destination
// line before overwriting it (read-for-ownership), which doubles the traffic of the part the CPU writes; the
// expansion itself (26 mask bytes from 26 bits) was ~20 scalar operations per env.  Here one 64-byte line of mask
// rows costs five instructions and goes out with a streaming store.
#include "skyjo_hostsimd.h"

#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <vector>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace skyjo {

constexpr int PACK_DONE_SH = 26, PACK_AGENT_SH = 28;

static bool mask_streaming_stores() {
    static const bool nt = [] {
        const char *g = getenv("SKYJO_HOST_NT_MASK");  // experiment knob
        return g && atoi(g) != 0;
    }();
    return nt;
}

// SKYJO_HOST_NT=0 (experiment knob): plain 64-byte stores instead of streaming ones for the record expansion
static bool use_streaming_stores() {
    static const bool nt = [] {
        const char *g = getenv("SKYJO_HOST_NT");
        return !(g && atoi(g) == 0);
    }();
    return nt;
}

// ---- portable ---------------------------------------------------------------------------------
static inline uint64_t spread8(uint32_t b) {
    const uint64_t x = ((uint64_t)(b & 0xFFu) * 0x0101010101010101ull) & 0x8040201008040201ull;
    return ((x + 0x7F7F7F7F7F7F7F7Full) >> 7) & 0x0101010101010101ull;
}

static void expand_packed_scalar(const uint32_t *packed, long long e0, long long e1, int8_t *mask, int8_t *agent,
                                 uint8_t *done) {
    for (long long e = e0; e < e1; ++e) {
        const uint32_t p = packed[e];
        const uint64_t m0 = spread8(p), m1 = spread8(p >> 8), m2 = spread8(p >> 16);
        const uint16_t m3 = (uint16_t)(spread8(p >> 24) & 0xFFFFu);
        int8_t *row = mask + e * 26;
        memcpy(row, &m0, 8);
        memcpy(row + 8, &m1, 8);
        memcpy(row + 16, &m2, 8);
        memcpy(row + 24, &m3, 2);
        agent[e] = (int8_t)(p >> PACK_AGENT_SH);
        done[e] = (uint8_t)((p >> PACK_DONE_SH) & 3u);
    }
}

#if defined(__x86_64__)
// ---- AVX-512 -----------------------------------------------------------------------------------
// Mask rows are 26 bytes, so 32 envs are 832 bytes = 13 lines of 64.  Output byte b of a 32-env group is bit
// b % 26 of env b / 26.  The bytes of line L come from at most 4 consecutive envs, i.e. from one 16-byte window
// of packed words starting at env floor(64 L / 26): broadcast the window to the four 128-bit lanes, pick for
// every output byte the source byte that holds its bit (pshufb), test the bit, and turn the 64 results into 64
// bytes of 0 / 1.
struct MaskTables {
    alignas(64) uint8_t idx[13][64];
    alignas(64) uint8_t sel[13][64];
    int base[13];
    MaskTables() {
        for (int L = 0; L < 13; ++L) {
            base[L] = (64 * L) / 26;
            for (int j = 0; j < 64; ++j) {
                const int b = 64 * L + j, env = b / 26, bit = b % 26;
                idx[L][j] = (uint8_t)(4 * (env - base[L]) + bit / 8);
                sel[L][j] = (uint8_t)(1u << (bit % 8));
            }
        }
    }
};

__attribute__((target("avx512f,avx512bw,avx512vl"))) static void expand_packed_avx512(
    const uint32_t *packed, long long e0, long long groups, int8_t *mask, int8_t *agent, uint8_t *done) {
    static const MaskTables T;
    // plain stores: measured on the B200 boxes, streaming the mask rows while the copy engine writes the observation
    // rows into the same host memory slows the copies down (raw wire, 4 workers: 5.3-5.8e8 env-steps/s streamed
    // against 7.3e8 plain)
    const bool nt = mask_streaming_stores();
    alignas(64) uint32_t local[64 + 16];
    for (long long g = 0; g < groups; ++g) {
        const uint32_t *src = packed + e0 + 64 * g;
        __m512i w[4];
        for (int q = 0; q < 4; ++q) {
            w[q] = _mm512_loadu_si512(src + 16 * q);
            _mm512_store_si512(local + 16 * q, w[q]);
        }
        _mm512_store_si512(local + 64, _mm512_setzero_si512());
        int8_t *mrow = mask + (e0 + 64 * g) * 26;
        for (int half = 0; half < 2; ++half) {
            const uint32_t *lw = local + 32 * half;
            int8_t *dst = mrow + 832 * half;
            for (int L = 0; L < 13; ++L) {
                const __m128i win = _mm_loadu_si128(reinterpret_cast<const __m128i *>(lw + T.base[L]));
                const __m512i bytes = _mm512_shuffle_epi8(_mm512_broadcast_i32x4(win),
                                                          _mm512_load_si512(T.idx[L]));
                const __mmask64 k = _mm512_test_epi8_mask(bytes, _mm512_load_si512(T.sel[L]));
                if (nt)
                    _mm512_stream_si512(reinterpret_cast<__m512i *>(dst + 64 * L), _mm512_maskz_set1_epi8(k, 1));
                else
                    _mm512_store_si512(reinterpret_cast<__m512i *>(dst + 64 * L), _mm512_maskz_set1_epi8(k, 1));
            }
        }
        __m512i ag = _mm512_setzero_si512(), dn = _mm512_setzero_si512();
        const __m512i three = _mm512_set1_epi32(3);
        ag = _mm512_inserti32x4(ag, _mm512_cvtepi32_epi8(_mm512_srli_epi32(w[0], PACK_AGENT_SH)), 0);
        ag = _mm512_inserti32x4(ag, _mm512_cvtepi32_epi8(_mm512_srli_epi32(w[1], PACK_AGENT_SH)), 1);
        ag = _mm512_inserti32x4(ag, _mm512_cvtepi32_epi8(_mm512_srli_epi32(w[2], PACK_AGENT_SH)), 2);
        ag = _mm512_inserti32x4(ag, _mm512_cvtepi32_epi8(_mm512_srli_epi32(w[3], PACK_AGENT_SH)), 3);
        dn = _mm512_inserti32x4(dn, _mm512_cvtepi32_epi8(_mm512_and_si512(_mm512_srli_epi32(w[0], PACK_DONE_SH), three)), 0);
        dn = _mm512_inserti32x4(dn, _mm512_cvtepi32_epi8(_mm512_and_si512(_mm512_srli_epi32(w[1], PACK_DONE_SH), three)), 1);
        dn = _mm512_inserti32x4(dn, _mm512_cvtepi32_epi8(_mm512_and_si512(_mm512_srli_epi32(w[2], PACK_DONE_SH), three)), 2);
        dn = _mm512_inserti32x4(dn, _mm512_cvtepi32_epi8(_mm512_and_si512(_mm512_srli_epi32(w[3], PACK_DONE_SH), three)), 3);
        if (nt) {
            _mm512_stream_si512(reinterpret_cast<__m512i *>(agent + e0 + 64 * g), ag);
            _mm512_stream_si512(reinterpret_cast<__m512i *>(done + e0 + 64 * g), dn);
        } else {
            _mm512_store_si512(reinterpret_cast<__m512i *>(agent + e0 + 64 * g), ag);
            _mm512_store_si512(reinterpret_cast<__m512i *>(done + e0 + 64 * g), dn);
        }
    }
    _mm_sfence();
}

// 8 record bytes (16 nibbles, low nibble first) -> 16 bytes, through a 16-entry table
__attribute__((target("ssse3"))) static inline __m128i nibbles16(const uint8_t *src, __m128i lut) {
    const __m128i x = _mm_loadl_epi64(reinterpret_cast<const __m128i *>(src));
    const __m128i m = _mm_set1_epi8(0x0F);
    const __m128i idx = _mm_unpacklo_epi8(_mm_and_si128(x, m), _mm_and_si128(_mm_srli_epi16(x, 4), m));
    return _mm_shuffle_epi8(lut, idx);
}

// one record -> one row at o (which may be followed by at least 16 writable bytes)
__attribute__((target("ssse3"))) static inline void expand_one_record(const uint8_t *r, int R, int8_t *o) {
    const int NC = 6 * R;
    const __m128i ident = _mm_setr_epi8(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15);
    const __m128i cardv = _mm_setr_epi8(-2, -1, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 15);
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
    if (i < NC) {  // 6 R is not a multiple of 8: the last 2 / 4 / 6 record bytes (the row's slack takes the overshoot)
        uint64_t tail = 0;
        memcpy(&tail, cards + i, (size_t)(NC - i));
        _mm_storeu_si128(reinterpret_cast<__m128i *>(o + 19 + 2 * i),
                         nibbles16(reinterpret_cast<const uint8_t *>(&tail), cardv));
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

__attribute__((target("avx512f,avx512bw,avx512vl,ssse3"))) static void expand_obs_avx512(
    const uint8_t *rec, long long e0, long long groups, int D, int RB, int8_t *obs) {
    const int R = (D - 19) / 12;
    alignas(64) int8_t stage[64 * (19 + 12 * 12) + 64];
    for (long long g = 0; g < groups; ++g) {
        const long long eb = e0 + 64 * g;
        // front to back: a row's 16-byte stores may overshoot into the next row's first bytes, which that row
        // then rewrites (the last row overshoots into the slack of `stage`)
        for (int k = 0; k < 64; ++k) expand_one_record(rec + (eb + k) * RB, R, stage + k * D);
        int8_t *dst = obs + eb * D;
        for (int L = 0; L < D; ++L)
            _mm512_stream_si512(reinterpret_cast<__m512i *>(dst + 64 * L), _mm512_load_si512(stage + 64 * L));
    }
    _mm_sfence();
}

// ---- AVX-512 VBMI: one output line per ~15 instructions, no staging ----------------------------------------
// The rows of 64 consecutive envs are 64 D bytes = D output lines.  Output byte b of the group belongs to row
// b / D, position j = b % D, and is a function of ONE byte of that env's record (csrc/skyjo_hostio.cuh has the
// record layout): a raw byte (j = 0, and the value-0 bin j = 4), or one of its nibbles taken as is (j = 1, the
// other bins), minus 3 (the discard top, j = 17) or through the card table (the hand j = 18 and the slots j >= 19;
// 15 stays 15 = hidden); a slot whose column is removed shows -14, which is one bit of the record's flag bytes.
// Per line the tables hold, for each of the 64 bytes, where its source byte and its flag byte sit inside two
// 128-byte windows of the group's records, which nibble, which value table, and the flag bit.  vpermt2b gathers,
// a 64-entry vpermb is the value table.
struct ObsLineTables {
    int D = 0, RB = 0;
    bool ok = false;
    std::vector<int> woff, foff;               // per line: byte offset of the two windows from the group's first record
    std::vector<uint8_t> sb, sbf, kind, fsel;  // per line x 64
    std::vector<uint64_t> k_hi, k_raw;         // per line: high-nibble positions, raw-byte positions
};

static void build_obs_tables(ObsLineTables &T, int D) {
    const int R = (D - 19) / 12, RB = 12 + 6 * R + (R + 1) / 2;
    T.D = D;
    T.RB = RB;
    T.ok = true;
    T.woff.assign(D, 0);
    T.foff.assign(D, 0);
    T.sb.assign((size_t)D * 64, 0);
    T.sbf.assign((size_t)D * 64, 0);
    T.kind.assign((size_t)D * 64, 0);
    T.fsel.assign((size_t)D * 64, 0);
    T.k_hi.assign(D, 0);
    T.k_raw.assign(D, 0);
    for (int L = 0; L < D; ++L) {
        int src[64], fsrc[64];
        int lo = 1 << 30, hi = -1, flo = 1 << 30, fhi = -1;
        for (int x = 0; x < 64; ++x) {
            const int b = 64 * L + x, row = b / D, j = b % D, base = row * RB;
            int byte = 0, nib_hi = 0, kind = 0, raw = 0, fb = -1, fbit = 0;  // kind: 0 as is, 1 minus 3, 2 card table
            if (j == 0) {
                byte = 0; raw = 1;
            } else if (j == 1) {
                byte = 1;
            } else if (j == 4) {
                byte = 3; raw = 1;
            } else if (j <= 16) {
                const int k = j - 2;
                byte = 4 + (k >> 1); nib_hi = k & 1;
            } else if (j == 17) {
                byte = 1; nib_hi = 1; kind = 1;
            } else if (j == 18) {
                byte = 2; kind = 2;
            } else {
                const int i = j - 19, q = i / 12, col = (i % 12) / 3;
                byte = 12 + (i >> 1); nib_hi = i & 1; kind = 2;
                fb = 12 + 6 * R + (q >> 1); fbit = 4 * (q & 1) + col;
            }
            src[x] = base + byte;
            fsrc[x] = fb < 0 ? -1 : base + fb;
            lo = src[x] < lo ? src[x] : lo;
            hi = src[x] > hi ? src[x] : hi;
            if (fb >= 0) {
                flo = fsrc[x] < flo ? fsrc[x] : flo;
                fhi = fsrc[x] > fhi ? fsrc[x] : fhi;
            }
            T.kind[(size_t)L * 64 + x] = (uint8_t)(16 * kind);
            T.fsel[(size_t)L * 64 + x] = (uint8_t)(fb < 0 ? 0 : (1 << fbit));
            if (nib_hi) T.k_hi[L] |= 1ull << x;
            if (raw) T.k_raw[L] |= 1ull << x;
        }
        if (fhi < 0) flo = fhi = lo;
        if (hi - lo >= 128 || fhi - flo >= 128) T.ok = false;
        T.woff[L] = lo;
        T.foff[L] = flo;
        for (int x = 0; x < 64; ++x) {
            T.sb[(size_t)L * 64 + x] = (uint8_t)((src[x] - lo) & 127);
            T.sbf[(size_t)L * 64 + x] = (uint8_t)(fsrc[x] < 0 ? 0 : ((fsrc[x] - flo) & 127));
        }
    }
}

__attribute__((target("avx512f,avx512bw,avx512vl,avx512vbmi"))) static void expand_obs_vbmi(
    const ObsLineTables &T, const uint8_t *rec, long long e0, long long groups, int8_t *obs) {
    const int D = T.D, RB = T.RB;
    alignas(64) int8_t vt[64];  // value tables: [0..15] as is, [16..31] minus 3, [32..47] cards (value + 2 -> value, 15 hidden)
    for (int i = 0; i < 16; ++i) {
        vt[i] = (int8_t)i;
        vt[16 + i] = (int8_t)(i - 3);
        vt[32 + i] = (int8_t)(i == 15 ? 15 : i - 2);
        vt[48 + i] = 0;
    }
    const bool nt = use_streaming_stores();
    const __m512i table = _mm512_load_si512(vt), low4 = _mm512_set1_epi8(0x0F), removed = _mm512_set1_epi8(-14);
    for (long long g = 0; g < groups; ++g) {
        const uint8_t *base = rec + (e0 + 64 * g) * RB;
        int8_t *dst = obs + (e0 + 64 * g) * D;
        for (int L = 0; L < D; ++L) {
            const uint8_t *w = base + T.woff[L], *f = base + T.foff[L];
            const __m512i t = _mm512_permutex2var_epi8(_mm512_loadu_si512(w), _mm512_loadu_si512(T.sb.data() + (size_t)L * 64),
                                                       _mm512_loadu_si512(w + 64));
            const __m512i n = _mm512_mask_blend_epi8(T.k_hi[L], _mm512_and_si512(t, low4),
                                                     _mm512_and_si512(_mm512_srli_epi16(t, 4), low4));
            __m512i v = _mm512_permutexvar_epi8(_mm512_or_si512(n, _mm512_loadu_si512(T.kind.data() + (size_t)L * 64)), table);
            v = _mm512_mask_blend_epi8(T.k_raw[L], v, t);
            const __m512i fl = _mm512_permutex2var_epi8(_mm512_loadu_si512(f), _mm512_loadu_si512(T.sbf.data() + (size_t)L * 64),
                                                        _mm512_loadu_si512(f + 64));
            const __mmask64 k = _mm512_test_epi8_mask(fl, _mm512_loadu_si512(T.fsel.data() + (size_t)L * 64));
            v = _mm512_mask_blend_epi8(k, v, removed);
            if (nt)
                _mm512_stream_si512(reinterpret_cast<__m512i *>(dst + 64 * L), v);
            else
                _mm512_store_si512(reinterpret_cast<__m512i *>(dst + 64 * L), v);
        }
    }
    _mm_sfence();
}

static const ObsLineTables *obs_tables_for(int D) {
    // one table set per row length (13 possible: D = 19 + 12 R), built on first use
    static ObsLineTables tables[13];
    static std::mutex m;
    const int R = (D - 19) / 12;
    if (R < 1 || R > 12 || D != 19 + 12 * R) return nullptr;
    std::lock_guard<std::mutex> g(m);
    if (tables[R].D != D) build_obs_tables(tables[R], D);
    return tables[R].ok ? &tables[R] : nullptr;
}
#endif

// ---- dispatch ------------------------------------------------------------------------------------
int host_simd_level() {
#if defined(__x86_64__)
    static const int level = (__builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") &&
                              __builtin_cpu_supports("avx512vl"))
                                 ? (__builtin_cpu_supports("avx512vbmi") ? 3 : 2) : 0;
    return level;
#else
    return 0;
#endif
}

void expand_packed_wide(const uint32_t *packed, long long e0, long long e1, int8_t *mask, int8_t *agent, uint8_t *done) {
#if defined(__x86_64__)
    if (host_simd_level() >= 2 && (((uintptr_t)mask | (uintptr_t)agent | (uintptr_t)done) & 63) == 0) {
        const long long a0 = (e0 + 63) / 64 * 64;  // first env of a 64-env group: 64 * 26 bytes and 64 bytes are line multiples
        if (a0 < e1) {
            const long long groups = (e1 - a0) / 64;
            expand_packed_scalar(packed, e0, a0, mask, agent, done);
            expand_packed_avx512(packed, a0, groups, mask, agent, done);
            expand_packed_scalar(packed, a0 + 64 * groups, e1, mask, agent, done);
            return;
        }
    }
#endif
    expand_packed_scalar(packed, e0, e1, mask, agent, done);
}

// portable record expansion lives in skyjo_hostio.cuh (expand_obs_records); declared here for the fallback
void expand_obs_records_portable(const uint8_t *rec, long long e0, long long e1, int D, int8_t *obs);

void expand_obs_records_wide(const uint8_t *rec, long long e0, long long e1, int D, int8_t *obs, long long rec_slack) {
#if defined(__x86_64__)
    if (host_simd_level() >= 2 && ((uintptr_t)obs & 63) == 0) {
        const int R = (D - 19) / 12, RB = 12 + 6 * R + (R + 1) / 2;
        const long long a0 = (e0 + 63) / 64 * 64;
        if (a0 < e1) {
            const long long groups = (e1 - a0) / 64;
            expand_obs_records_portable(rec, e0, a0, D, obs);
            // the gathers read 128-byte windows of the record array: the last group of the whole array may not
            // (the caller's array ends there), so it takes the staged version
            const ObsLineTables *T = host_simd_level() >= 3 && rec_slack >= 256 ? obs_tables_for(D) : nullptr;
            if (T)
                expand_obs_vbmi(*T, rec, a0, groups, obs);
            else
                expand_obs_avx512(rec, a0, groups, D, RB, obs);
            expand_obs_records_portable(rec, a0 + 64 * groups, e1, D, obs);
            return;
        }
    }
#endif
    expand_obs_records_portable(rec, e0, e1, D, obs);
}

}  // namespace skyjo
