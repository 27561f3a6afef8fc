/*
 * fdf_avx2_port.cpp -- AVX2 CPU port of the reference's hot path, used ONLY as the timed CPU
 * baseline ("cpu_baseline.kind": "port") and as a second checker for the scalar oracle.
 *
 * TEST / BENCH INFRASTRUCTURE ONLY (see fdf_oracle.h).  The reference crate is Rust and cannot be
 * built here (no cargo/rustc), so the same algorithm is restated with the same x86 instruction mix
 * in C++ <immintrin.h>:
 *   - row scan in blocks of 16 centres + cardinal 2-of-4 / 3-of-4 pre-check  fast_simd.rs:368-520
 *   - per-candidate two 8-lane dword gathers + byte shuffles -> 16 ring bytes fast_simd.rs:132-216
 *   - rotating n-byte mask segment test with ptest                           fast_simd.rs:218-297
 *   - MaxThreshold score on 16 u16 lanes with phminposuw                     fast_simd.rs:623-718
 *   - SumAbsolute score with psadbw                                          fast_simd.rs:722-749
 *   - three rolling u16 score rows + deferred 3x3 strict-max NMS             fast_simd.rs:310-366, 588-616
 *   - row tail without pre-check                                             fast_simd.rs:558-586
 * It is a restatement of the reference, NOT the reference binary; it is validated bit-exact
 * against the scalar oracle (fdf_oracle.c) by tests/test_oracle.py before it is timed.
 *
 * The dword gathers read up to 3 bytes past each ring pixel: callers must pad the
 * image buffer by >= 4 bytes (oracle/oracle.py does).
 */
#include <immintrin.h>

#include <cstddef>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "fdf_oracle.h"

namespace {

constexpr int kRing[16][2] = {
    {0, -3}, {1, -3}, {2, -2}, {3, -1}, {3, 0},  {3, 1},   {2, 2},   {1, 3},
    {0, 3},  {-1, 3}, {-2, 2}, {-3, 1}, {-3, 0}, {-3, -1}, {-2, -2}, {-1, -3},
};
constexpr int kNorth = 0, kEast = 4, kSouth = 8, kWest = 12;

struct RingOffsets {
    alignas(32) int32_t at[16];
    explicit RingOffsets(uint32_t pitch) {
        for (int i = 0; i < 16; i++) at[i] = kRing[i][1] * (int32_t)pitch + kRing[i][0];
    }
};

inline __m128i gt_u8(__m128i a, __m128i b) {
    const __m128i bias = _mm_set1_epi8((char)0x80);
    return _mm_cmpgt_epi8(_mm_xor_si128(a, bias), _mm_xor_si128(b, bias));
}

inline __m128i first_n_bytes(int n) {
    alignas(16) uint8_t m[16] = {0};
    for (int i = 0; i < n; i++) m[i] = 0xff;
    return _mm_load_si128((const __m128i *)m);
}

inline __m256i first_n_words(int n) {
    alignas(32) uint16_t m[16] = {0};
    for (int i = 0; i < n; i++) m[i] = 0xffff;
    return _mm256_load_si256((const __m256i *)m);
}

inline uint16_t hmin_u16(__m256i v) {
    __m128i lo = _mm_minpos_epu16(_mm256_castsi256_si128(v));
    __m128i hi = _mm_minpos_epu16(_mm256_extracti128_si256(v, 1));
    uint16_t a = (uint16_t)_mm_extract_epi16(lo, 0), b = (uint16_t)_mm_extract_epi16(hi, 0);
    return a < b ? a : b;
}

/* rotate the 16 u16 lanes down by one (lane i <- lane i+1), crossing the 128-bit halves */
inline __m256i rot1_u16(__m256i v) {
    __m256i in_lane = _mm256_alignr_epi8(v, v, 2);
    __m256i swapped = _mm256_permute2x128_si256(in_lane, in_lane, 1);
    const __m256i top = _mm256_set_epi64x((long long)0xFFFF000000000000ULL, 0,
                                          (long long)0xFFFF000000000000ULL, 0);
    return _mm256_blendv_epi8(in_lane, swapped, top);
}

inline uint32_t hsum_u8(__m128i v) {
    __m128i s = _mm_sad_epu8(v, _mm_setzero_si128());
    return (uint32_t)_mm_cvtsi128_si32(s) + (uint32_t)_mm_extract_epi16(s, 4);
}

/* fast_simd.rs:623-718 */
uint16_t score_max_threshold(uint8_t centre, __m128i ring, int n) {
    __m256i wide = _mm256_cvtepu8_epi16(ring); /* lanes 0..15 = ring 0..15 */
    __m256i diff = _mm256_sub_epi16(_mm256_set1_epi16((short)(centre + 512)), wide);
    const __m256i sel = first_n_words(n);
    const __m256i fill = _mm256_andnot_si256(sel, _mm256_set1_epi8(-1));
    const __m256i ones = _mm256_set1_epi16(-1);
    alignas(32) uint16_t wmin[16], wmax_inv[16];
    for (int k = 0; k < 16; k++) {
        wmin[k] = hmin_u16(_mm256_or_si256(_mm256_and_si256(diff, sel), fill));
        __m256i inv = _mm256_sub_epi16(ones, diff);
        wmax_inv[k] = hmin_u16(_mm256_or_si256(_mm256_and_si256(inv, sel), fill));
        diff = rot1_u16(diff);
    }
    const __m256i k1024 = _mm256_set1_epi16(1024);
    uint16_t a = hmin_u16(_mm256_sub_epi16(k1024, _mm256_load_si256((const __m256i *)wmin)));
    int extreme_highest = 1024 - (int)a - 512;
    uint16_t b = hmin_u16(_mm256_sub_epi16(k1024, _mm256_load_si256((const __m256i *)wmax_inv)));
    int extreme_lowest = (int)(int16_t)b - (1024 + 512 + 1);
    extreme_lowest = (int)(int16_t)extreme_lowest;
    int ah = extreme_highest < 0 ? -extreme_highest : extreme_highest;
    int al = extreme_lowest < 0 ? -extreme_lowest : extreme_lowest;
    return (uint16_t)(ah < al ? ah : al);
}

/* fast_simd.rs:722-749 */
inline uint16_t score_sum_abs(__m128i ring, __m128i centre, __m128i above, __m128i below, __m128i thr) {
    __m128i c_minus_p = _mm_and_si128(_mm_subs_epu8(_mm_subs_epu8(centre, ring), thr), below);
    __m128i p_minus_c = _mm_and_si128(_mm_subs_epu8(_mm_subs_epu8(ring, centre), thr), above);
    uint32_t s0 = hsum_u8(c_minus_p), s1 = hsum_u8(p_minus_c);
    return (uint16_t)(s0 > s1 ? s0 : s1);
}

/* fast_simd.rs:115-297 */
template <int MODE>
inline bool full_test(const uint8_t *centre_ptr, const RingOffsets &off, uint8_t t, int n,
                      __m128i run_mask, uint16_t *score) {
    const uint8_t c = *centre_ptr;
    const int *base = (const int *)centre_ptr;
    const __m256i pick_lo = _mm256_set_epi64x((long long)0x8080808080808080ULL, (long long)0x808080800c080400ULL,
                                              (long long)0x8080808080808080ULL, (long long)0x808080800c080400ULL);
    const __m256i pick_hi = _mm256_set_epi64x((long long)0x808080800c080400ULL, (long long)0x8080808080808080ULL,
                                              (long long)0x808080800c080400ULL, (long long)0x8080808080808080ULL);
    __m256i g0 = _mm256_i32gather_epi32(base, _mm256_load_si256((const __m256i *)&off.at[0]), 1);
    __m256i g1 = _mm256_i32gather_epi32(base, _mm256_load_si256((const __m256i *)&off.at[8]), 1);
    __m256i packed = _mm256_or_si256(_mm256_shuffle_epi8(g0, pick_lo), _mm256_shuffle_epi8(g1, pick_hi));
    /* dwords: [r0-3, 0, r8-11, 0 | r4-7, 0, r12-15, 0] -> [r0-3, r4-7, r8-11, r12-15 | ...] */
    const __m256i order = _mm256_set_epi32(1, 1, 1, 1, 6, 2, 4, 0);
    __m128i ring = _mm256_castsi256_si128(_mm256_permutevar8x32_epi32(packed, order));

    const __m128i vc = _mm_set1_epi8((char)c), vt = _mm_set1_epi8((char)t);
    const __m128i above = gt_u8(ring, _mm_adds_epu8(vc, vt));
    const __m128i below = gt_u8(_mm_subs_epu8(vc, vt), ring);
    const __m128i all = _mm_set1_epi8(-1);
    for (int k = 0; k < 16; k++) {
        __m128i outside = _mm_andnot_si128(run_mask, all);
        if (_mm_test_all_ones(_mm_or_si128(_mm_and_si128(above, run_mask), outside)) ||
            _mm_test_all_ones(_mm_or_si128(_mm_and_si128(below, run_mask), outside))) {
            if (MODE == FDF_ORACLE_NMS_MAX_THRESHOLD) *score = score_max_threshold(c, ring, n);
            if (MODE == FDF_ORACLE_NMS_SUM_ABSOLUTE) *score = score_sum_abs(ring, vc, above, below, vt);
            return true;
        }
        run_mask = _mm_alignr_epi8(run_mask, run_mask, 1);
    }
    return false;
}

/* fast_simd.rs:301-620 */
template <int MODE>
int64_t run(const uint8_t *img, uint32_t w, uint32_t h, uint32_t pitch, uint8_t t, int n,
            fdf_oracle_point *out, size_t cap) {
    const RingOffsets off(pitch);
    const __m128i vt = _mm_set1_epi8((char)t);
    const __m128i run_mask = first_n_bytes(n);
    std::vector<uint16_t> rows((size_t)w * 3, 0);
    uint16_t *row_buf[3] = {rows.data(), rows.data() + w, rows.data() + 2 * (size_t)w};
    int64_t count = 0;
    auto emit = [&](uint32_t x, uint32_t y) {
        if ((size_t)count < cap) {
            out[count].x = x;
            out[count].y = y;
        }
        count++;
    };
    const uint32_t blocks = (w - 6) / 16;

    for (uint32_t y = 3; y < h - 3; y++) {
        uint16_t *above = row_buf[(y + 0) % 3], *centre = row_buf[(y + 1) % 3], *cur = row_buf[(y + 2) % 3];
        if (MODE != FDF_ORACLE_NMS_OFF) std::memset(cur, 0, (size_t)w * sizeof(uint16_t));
        const uint8_t *row = img + (size_t)y * pitch;

        for (uint32_t b = 0; b < blocks; b++) {
            const uint32_t x0 = 3 + b * 16;
            const uint8_t *p = row + x0;
            __m128i c = _mm_loadu_si128((const __m128i *)p);
            __m128i north = _mm_loadu_si128((const __m128i *)(p + off.at[kNorth]));
            __m128i east = _mm_loadu_si128((const __m128i *)(p + off.at[kEast]));
            __m128i south = _mm_loadu_si128((const __m128i *)(p + off.at[kSouth]));
            __m128i west = _mm_loadu_si128((const __m128i *)(p + off.at[kWest]));
            __m128i hi = _mm_adds_epu8(c, vt), lo = _mm_subs_epu8(c, vt);
            __m128i na = gt_u8(north, hi), ea = gt_u8(east, hi), sa = gt_u8(south, hi), wa = gt_u8(west, hi);
            __m128i nb = gt_u8(lo, north), eb = gt_u8(lo, east), sb = gt_u8(lo, south), wb = gt_u8(lo, west);
            __m128i cand;
            if (n < 12) { /* two adjacent cardinals */
                __m128i a = _mm_or_si128(_mm_or_si128(_mm_and_si128(sa, wa), _mm_and_si128(na, wa)),
                                         _mm_or_si128(_mm_and_si128(na, ea), _mm_and_si128(ea, sa)));
                __m128i d = _mm_or_si128(_mm_or_si128(_mm_and_si128(sb, wb), _mm_and_si128(nb, wb)),
                                         _mm_or_si128(_mm_and_si128(nb, eb), _mm_and_si128(eb, sb)));
                cand = _mm_or_si128(a, d);
            } else { /* three of four cardinals */
                __m128i a = _mm_or_si128(
                    _mm_or_si128(_mm_and_si128(_mm_and_si128(ea, sa), wa), _mm_and_si128(_mm_and_si128(na, sa), wa)),
                    _mm_or_si128(_mm_and_si128(_mm_and_si128(na, ea), wa), _mm_and_si128(_mm_and_si128(na, ea), sa)));
                __m128i d = _mm_or_si128(
                    _mm_or_si128(_mm_and_si128(_mm_and_si128(eb, sb), wb), _mm_and_si128(_mm_and_si128(nb, sb), wb)),
                    _mm_or_si128(_mm_and_si128(_mm_and_si128(nb, eb), wb), _mm_and_si128(_mm_and_si128(nb, eb), sb)));
                cand = _mm_or_si128(a, d);
            }
            if (_mm_test_all_zeros(cand, cand)) continue;
            __m128i probe = _mm_set_epi64x(0, 0xff);
            for (uint32_t x = x0; x < x0 + 16; x++) {
                bool skip = _mm_test_all_zeros(cand, probe);
                probe = _mm_bslli_si128(probe, 1);
                if (skip) continue;
                uint16_t s = 0;
                if (full_test<MODE>(row + x, off, t, n, run_mask, &s)) {
                    if (MODE == FDF_ORACLE_NMS_OFF) emit(x, y);
                    else cur[x] = s;
                }
            }
        }
        for (uint32_t x = 3 + blocks * 16; x < w - 3; x++) { /* tail */
            uint16_t s = 0;
            if (full_test<MODE>(row + x, off, t, n, run_mask, &s)) {
                if (MODE == FDF_ORACLE_NMS_OFF) emit(x, y);
                else cur[x] = s;
            }
        }
        if (MODE != FDF_ORACLE_NMS_OFF) {
            if (y == 4) continue; /* row 3 is never finalised (fast_simd.rs:590-592) */
            for (uint32_t x = 3; x < w - 3; x++) {
                const uint16_t s = centre[x];
                if (s == 0) continue;
                if (s > above[x - 1] && s > above[x] && s > above[x + 1] && s > centre[x - 1] &&
                    s > centre[x + 1] && s > cur[x - 1] && s > cur[x] && s > cur[x + 1])
                    emit(x, y - 1);
            }
        }
    }
    return count;
}

}  // namespace

extern "C" {

/* Same contract as fdf_oracle_detect (without scores).  img must be padded by >= 4 bytes. */
int64_t fdf_avx2_port_detect(const uint8_t *img, uint32_t w, uint32_t h, uint32_t pitch, uint8_t t,
                             uint8_t n, uint8_t nms, fdf_oracle_point *out, size_t cap) {
    if (n < 9 || n > 16) return -1;
    if (w < 7 || h < 7) return 0;
    switch (nms) {
        case FDF_ORACLE_NMS_OFF: return run<FDF_ORACLE_NMS_OFF>(img, w, h, pitch, t, n, out, cap);
        case FDF_ORACLE_NMS_MAX_THRESHOLD: return run<FDF_ORACLE_NMS_MAX_THRESHOLD>(img, w, h, pitch, t, n, out, cap);
        case FDF_ORACLE_NMS_SUM_ABSOLUTE: return run<FDF_ORACLE_NMS_SUM_ABSOLUTE>(img, w, h, pitch, t, n, out, cap);
        default: return -2;
    }
}

/*
 * "One frame per core": n_threads workers, worker k takes frames k, k+n_threads, ...  Each frame is
 * detected independently exactly as fdf_avx2_port_detect does; counts[f] receives the keypoint count
 * and hashes[f] (optional) the tests/compare.rs-style SipHash of the frame's point list.  The frame
 * buffer must be padded by >= 4 bytes after the last frame.  Used only for the CPU baseline timing.
 */
int fdf_avx2_port_detect_batch(const uint8_t *frames, uint32_t n_frames, uint32_t w, uint32_t h,
                               uint32_t pitch, uint64_t frame_stride, uint8_t t, uint8_t n, uint8_t nms,
                               int64_t *counts, uint64_t *hashes, uint32_t n_threads) {
    if (n < 9 || n > 16) return -1;
    if (nms > FDF_ORACLE_NMS_SUM_ABSOLUTE) return -2;
    if (n_threads == 0) n_threads = 1;
    auto worker = [&](uint32_t k) {
        std::vector<fdf_oracle_point> pts((size_t)w * h / 64 + 16);
        for (uint32_t f = k; f < n_frames; f += n_threads) {
            const uint8_t *img = frames + (size_t)f * frame_stride;
            int64_t c = fdf_avx2_port_detect(img, w, h, pitch, t, n, nms, pts.data(), pts.size());
            if (c > (int64_t)pts.size()) {
                pts.resize((size_t)c);
                c = fdf_avx2_port_detect(img, w, h, pitch, t, n, nms, pts.data(), pts.size());
            }
            counts[f] = c;
            if (hashes) hashes[f] = fdf_oracle_hash_points(pts.data(), (size_t)c);
        }
    };
    if (n_threads == 1) {
        worker(0);
        return 0;
    }
    std::vector<std::thread> pool;
    for (uint32_t k = 0; k < n_threads; k++) pool.emplace_back(worker, k);
    for (auto &th : pool) th.join();
    return 0;
}

/* Score helpers on raw (centre, ring) vectors, for the reference's randomised equality tests
 * (fast_simd.rs:919-948 and :1185-1236). */
uint16_t fdf_avx2_port_score_max_threshold(uint8_t centre, const uint8_t ring[16], uint8_t n) {
    return score_max_threshold(centre, _mm_loadu_si128((const __m128i *)ring), n);
}

uint16_t fdf_avx2_port_score_sum_abs(uint8_t centre, const uint8_t ring[16], uint8_t t) {
    __m128i r = _mm_loadu_si128((const __m128i *)ring);
    __m128i vc = _mm_set1_epi8((char)centre), vt = _mm_set1_epi8((char)t);
    __m128i above = gt_u8(r, _mm_adds_epu8(vc, vt));
    __m128i below = gt_u8(_mm_subs_epu8(vc, vt), r);
    return score_sum_abs(r, vc, above, below, vt);
}

/*
 * The reference's randomised score tests (port == scalar on random vectors) with the same family
 * of generator: Xoshiro256++ seeded through SplitMix64, draw order as in the reference.  The
 * rand_xoshiro crate is not vendored, so bit-identical streams are not claimed; the property
 * tested (SIMD form == scalar form for arbitrary pixels) is the reference's.  Each function
 * returns the number of mismatches between the AVX2 port and the scalar oracle.
 */
struct Xoshiro256pp {
    uint64_t s[4];
    explicit Xoshiro256pp(uint64_t seed) {
        for (int i = 0; i < 4; i++) {
            seed += 0x9E3779B97F4A7C15ULL;
            uint64_t z = seed;
            z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
            z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
            s[i] = z ^ (z >> 31);
        }
    }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next_u64() {
        uint64_t result = rotl(s[0] + s[3], 23) + s[0];
        uint64_t t = s[1] << 17;
        s[2] ^= s[0];
        s[3] ^= s[1];
        s[1] ^= s[2];
        s[0] ^= s[3];
        s[2] ^= t;
        s[3] = rotl(s[3], 45);
        return result;
    }
    uint32_t next_u32() { return (uint32_t)(next_u64() >> 32); }
};

/* fast_simd.rs:883-894 + :939-945: per seed i: centre, then 16 ring values; count fixed at n. */
int64_t fdf_kat_random_max_threshold(uint64_t n_seeds, uint8_t n) {
    int64_t bad = 0;
    for (uint64_t i = 0; i < n_seeds; i++) {
        Xoshiro256pp rng(i);
        uint8_t centre = (uint8_t)rng.next_u32();
        uint8_t ring[16];
        for (int k = 0; k < 16; k++) ring[k] = (uint8_t)rng.next_u32();
        if (fdf_avx2_port_score_max_threshold(centre, ring, n) != fdf_oracle_score_max_threshold_px(centre, ring, n))
            bad++;
    }
    return bad;
}

/* fast_simd.rs:1198-1236: seed 0; per iteration: 16 ring bytes, centre, t. */
int64_t fdf_kat_random_sum_abs(uint64_t iterations) {
    int64_t bad = 0;
    Xoshiro256pp rng(0);
    for (uint64_t i = 0; i < iterations; i++) {
        uint8_t ring[16];
        for (int k = 0; k < 16; k++) ring[k] = (uint8_t)rng.next_u32();
        uint8_t centre = (uint8_t)rng.next_u32();
        uint8_t t = (uint8_t)rng.next_u32();
        if (fdf_avx2_port_score_sum_abs(centre, ring, t) != fdf_oracle_score_sum_abs_px(centre, ring, t)) bad++;
    }
    return bad;
}

}  // extern "C"
