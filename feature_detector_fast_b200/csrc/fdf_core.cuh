// fdf_core.cuh -- pixel-level primitives of the FAST-n path (segment test, scores, SWAR filter).
//
// Everything here is `__host__ __device__` so the exact device arithmetic can also be compiled
// with g++ and checked against the CPU oracle without a GPU (tests/host/core_check.cpp).  On the
// device each helper maps to one native sm_100a instruction where one exists
// (VABSDIFF4.U8[.ACC], PRMT, SHF.L.W, VIMNMX3, POPC).
//
// Reference semantics (citations into the reference checkout):
//   ring order / offsets     src/fast_simd.rs:79-98
//   brighter / darker        src/fast_simd.rs:218-231   (strict: p > c+t, p < c-t)
//   segment test             src/fast_simd.rs:244-296   (exists a cyclic run of >= n)
//   cardinal pre-check       src/fast_simd.rs:441-509   (a necessary-condition filter only)
//   MaxThreshold score       src/fast_simd.rs:623-718 == src/opencv_compat.rs:172-209
//   SumAbsolute score        src/fast_simd.rs:722-749 == src/opencv_compat.rs:278-299
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define FDF_HD __host__ __device__ __forceinline__
#else
#define FDF_HD inline
#endif

#if !defined(__CUDACC__)
#include <algorithm>
#endif

namespace fdf {

#if !defined(__CUDACC__)
using std::max;
using std::min;
#endif

enum : int { NMS_OFF = 0, NMS_MAX_THRESHOLD = 1, NMS_SUM_ABSOLUTE = 2 };

// Ring offsets (dx, dy), index 0 = north, clockwise, y down.  fast_simd.rs:79-98
#define FDF_RING_DX(i) ((i) == 0 ? 0 : (i) == 1 ? 1 : (i) == 2 ? 2 : (i) == 3 ? 3 : (i) == 4 ? 3 : (i) == 5 ? 3 : (i) == 6 ? 2 : (i) == 7 ? 1 : (i) == 8 ? 0 : (i) == 9 ? -1 : (i) == 10 ? -2 : (i) == 11 ? -3 : (i) == 12 ? -3 : (i) == 13 ? -3 : (i) == 14 ? -2 : -1)
#define FDF_RING_DY(i) ((i) == 0 ? -3 : (i) == 1 ? -3 : (i) == 2 ? -2 : (i) == 3 ? -1 : (i) == 4 ? 0 : (i) == 5 ? 1 : (i) == 6 ? 2 : (i) == 7 ? 3 : (i) == 8 ? 3 : (i) == 9 ? 3 : (i) == 10 ? 2 : (i) == 11 ? 1 : (i) == 12 ? 0 : (i) == 13 ? -1 : (i) == 14 ? -2 : -3)

// ---- one-instruction helpers -------------------------------------------------------------
FDF_HD uint32_t absdiff4(uint32_t a, uint32_t b) {  // per-byte |a - b|          VABSDIFF4.U8
#if defined(__CUDA_ARCH__)
    return __vabsdiffu4(a, b);
#else
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) {
        int x = (a >> (8 * i)) & 0xff, y = (b >> (8 * i)) & 0xff;
        r |= (uint32_t)(x > y ? x - y : y - x) << (8 * i);
    }
    return r;
#endif
}

FDF_HD uint32_t byte_perm(uint32_t a, uint32_t b, uint32_t sel) {  // PRMT
#if defined(__CUDA_ARCH__)
    return __byte_perm(a, b, sel);
#else
    uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) r |= (uint32_t)((v >> (8 * ((sel >> (4 * i)) & 7))) & 0xff) << (8 * i);
    return r;
#endif
}

FDF_HD uint32_t shift_in_sign(uint32_t acc, int v) {  // (acc << 1) | (v < 0)    SHF.L.W.U32.HI
#if defined(__CUDA_ARCH__)
    return __funnelshift_l((uint32_t)v, acc, 1);
#else
    return (acc << 1) | ((uint32_t)v >> 31);
#endif
}

FDF_HD int popc32(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}

// Three-input min / max: ptxas fuses two dependent 2-input ops into one VIMNMX3 when the inner result has a
// single use; the inner op goes through inline PTX so the front end cannot re-associate it for reuse
// elsewhere (which would turn 16 VIMNMX3 of a sliding window into 32 two-input ops).
FDF_HD int min3i(int a, int b, int c) {
#if defined(__CUDA_ARCH__)
    int m;
    asm("min.s32 %0, %1, %2;" : "=r"(m) : "r"(b), "r"(c));
    return min(a, m);
#else
    return min(a, min(b, c));
#endif
}
FDF_HD int max3i(int a, int b, int c) {
#if defined(__CUDA_ARCH__)
    int m;
    asm("max.s32 %0, %1, %2;" : "=r"(m) : "r"(b), "r"(c));
    return max(a, m);
#else
    return max(a, max(b, c));
#endif
}

// ---- dense SWAR filter (4 horizontally adjacent centres per 32-bit word) ---------------------
//
// A run of >= 9 of the 16 ring positions always contains at least one pixel of every
// diametrically opposite pair (the complement has <= 7 positions, the pair is 8 apart), and every
// pixel of the run differs from the centre by more than t.  So for the pairs north/south (ring 0/8)
// and east/west (ring 4/12):   keypoint  =>  max(|N-c|,|S-c|) > t  and  max(|E-c|,|W-c|) > t.
// max(a,b) > t implies (a|b) > t, which is what is tested: one OR instead of a byte-wise max.
// This is the same kind of necessary-condition filter as the reference's 2-of-4 / 3-of-4 cardinal
// pre-check (fast_simd.rs:441-509): it never changes the result, only which centres get the full
// test.  It is direction-less and slightly looser; it holds for every n in 9..=16.
//
// kbias = (0x7f - t) * 0x01010101 for t < 128, and 0 for t >= 128 (then only bit 7 of the
// difference is tested: |d| > t >= 128 implies |d| >= 128).  Per byte: bit 7 of
// ((m & 0x7f) + kbias) | m is set  <=>  m > t (t < 128)  or  m >= 128 (t >= 128).
FDF_HD uint32_t filter_kbias(uint32_t t) { return t < 128u ? (0x7fu - t) * 0x01010101u : 0u; }

FDF_HD uint32_t pair_exceeds(uint32_t a, uint32_t b, uint32_t c, uint32_t kbias) {
    const uint32_t da = absdiff4(a, c), db = absdiff4(b, c);
    const uint32_t x = ((da | db) & 0x7f7f7f7fu) + kbias;  // LOP3, IADD
    return x | da | db;                                    // LOP3; bit 7 of each byte is the flag
}

// Returns candidate flags in bit 7 of each byte, already ANDed with `valid` (0x80 per byte that
// is allowed to be a centre at all).
FDF_HD uint32_t filter4(uint32_t c, uint32_t n, uint32_t s, uint32_t e, uint32_t w, uint32_t kbias,
                        uint32_t valid) {
    return pair_exceeds(n, s, c, kbias) & pair_exceeds(e, w, c, kbias) & valid;
}

// ---- exact segment test ------------------------------------------------------------------------
struct RingMasks {
    uint32_t bright;  // bit i set <=> ring[i] > c + t   (fast_simd.rs:224 is_above)
    uint32_t dark;    // bit i set <=> ring[i] < c - t   (fast_simd.rs:225 is_below)
};

FDF_HD RingMasks ring_masks(int c, const int ring[16], int t) {
    const int hi = c + t, lo = c - t;
    // four independent 8-step shift chains (two per mask) instead of two 16-step ones: shorter critical path
    uint32_t bl = 0u, bh = 0u, dl = 0u, dh = 0u;
#pragma unroll
    for (int i = 7; i >= 0; i--) {
        bl = shift_in_sign(bl, hi - ring[i]);      // hi - p < 0  <=>  p > c + t
        bh = shift_in_sign(bh, hi - ring[i + 8]);
        dl = shift_in_sign(dl, ring[i] - lo);      // p - lo < 0  <=>  p < c - t
        dh = shift_in_sign(dh, ring[i + 8] - lo);
    }
    RingMasks m;
    m.bright = bl | (bh << 8);
    m.dark = dl | (dh << 8);
    return m;
}

// exists a cyclic run of >= n set bits in the 16-bit ring mask (9 <= n <= 16)
FDF_HD bool has_arc(uint32_t m16, int n) {
    uint32_t r = m16 | (m16 << 16);
    r &= r >> 1;
    r &= r >> 2;
    r &= r >> 4;         // bit i: positions i..i+7 all set
    r &= r >> (n - 8);   // bit i: positions i..i+n-1 all set
    return (r & 0xffffu) != 0u;
}

// ---- scores --------------------------------------------------------------------------------------
//
// MaxThreshold (opencv_compat.rs:172-209): with d_i = c - p_i and W_k the cyclic window of n ring
// positions starting at k,  eh = max_k min_{W_k} d,  el = min_k max_{W_k} d,  score = min(|eh|,|el|).
// Any two windows of >= 9 of 16 positions overlap, hence eh <= el.  For a keypoint whose arc is
// brighter than the centre (p > c+t on the arc) el <= -(t+1) < 0, so eh <= el < 0 and the score is
// -el = max_k min_{W_k} (p - c); for an arc darker than the centre (p < c-t) eh >= t+1 > 0, so
// el >= eh > 0 and the score is eh = max_k min_{W_k} (c - p).  So for keypoints
// (the only pixels that are ever scored, fast_simd.rs:276-279) one sliding-window max-of-min over
// e_i = +-(c - p_i) is exact.  Window minima: 3-window, then 9-window = min3 of three 3-windows,
// then n-window = min(9-window at k, 9-window at k+n-9).
template <int K>
FDF_HD int max_of_extended(const int u9[16]) {
    int v[16];
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = (K == 0) ? u9[i] : min(u9[i], u9[(i + K) & 15]);
    int a = max3i(v[0], v[1], v[2]), b = max3i(v[3], v[4], v[5]), c = max3i(v[6], v[7], v[8]);
    int d = max3i(v[9], v[10], v[11]), e = max3i(v[12], v[13], v[14]);
    return max(max3i(a, b, c), max3i(d, e, v[15]));
}

// pixel_is_brighter: the arc found by the segment test is a "bright" arc (ring pixels > c + t)
FDF_HD uint32_t score_max_threshold(int c, const int ring[16], int n, bool pixel_is_brighter) {
    int e[16], t3[16], u9[16];
#pragma unroll
    for (int i = 0; i < 16; i++) e[i] = pixel_is_brighter ? ring[i] - c : c - ring[i];
#pragma unroll
    for (int i = 0; i < 16; i++) t3[i] = min3i(e[i], e[(i + 1) & 15], e[(i + 2) & 15]);
#pragma unroll
    for (int i = 0; i < 16; i++) u9[i] = min3i(t3[i], t3[(i + 3) & 15], t3[(i + 6) & 15]);
    int r;
    switch (n) {
        case 9: r = max_of_extended<0>(u9); break;
        case 10: r = max_of_extended<1>(u9); break;
        case 11: r = max_of_extended<2>(u9); break;
        case 12: r = max_of_extended<3>(u9); break;
        case 13: r = max_of_extended<4>(u9); break;
        case 14: r = max_of_extended<5>(u9); break;
        case 15: r = max_of_extended<6>(u9); break;
        default: r = max_of_extended<7>(u9); break;
    }
    return (uint32_t)r;
}

// SumAbsolute (opencv_compat.rs:278-299): max( sum_{p > c+t} (p-c-t), sum_{p < c-t} (c-p-t) ) over
// ALL 16 ring pixels.  p - c - t > 0 <=> p > c + t, so each term is a relu.
FDF_HD uint32_t score_sum_abs(int c, const int ring[16], int t) {
    const int hi = c + t, lo = c - t;
    int sum_bright = 0, sum_dark = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        sum_bright += max(ring[i] - hi, 0);
        sum_dark += max(lo - ring[i], 0);
    }
    return (uint32_t)max(sum_bright, sum_dark);
}

}  // namespace fdf
