// fdf_core.cuh -- pixel-level primitives of the FAST-n path (SWAR filter, segment test, scores).
//
// Everything here is `__host__ __device__` so the exact device arithmetic can also be compiled
// with g++ and checked against the CPU oracle without a GPU (tests/host/strip_emulator.cpp).
//
// What the hardware dictates (measured with tools/pipe_probe.cu on a B200): LOP3, VABSDIFF4, PRMT, SHF,
// VIMNMX* and ISETP all issue on ONE pipe at 0.5 warp-instructions / clock / scheduler; IMAD (and adds /
// left shifts expressed as IMAD) issue on a second pipe at the same rate; POPC / FLO run at 1/8.  The path is
// bound by the first pipe, so the code below (a) works on 4 pixels per instruction in the dense filter,
// (b) works on 2 ring pixels per instruction (16-bit lanes) in the per-candidate test and scores, and
// (c) phrases adds, doublings and packing as multiply-adds wherever that is free.
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

// Index checks of the shared-memory structures (compute-sanitizer is closed on the development pool, so the kernels
// carry their own): FDF_BOUND(i, n) records the source line when i is outside [0, n).  Always active in the host
// build (tests/host/strip_emulator.cpp runs every CPU-tier image through them); on the device only in -DFDF_CHECKS
// builds (tools/checks_build.sh), where fdf_debug_check_failure() reads the line back (0 = clean).
#if defined(__CUDACC__) && defined(FDF_CHECKS)
__device__ int g_check_failure_line;
#endif
#if !defined(__CUDA_ARCH__)
inline int &check_failure_line() {
    static int line = 0;
    return line;
}
#define FDF_BOUND(i, n)                                                                       \
    do {                                                                                      \
        if ((unsigned long long)(i) >= (unsigned long long)(n)) ::fdf::check_failure_line() = __LINE__; \
    } while (0)
#elif defined(FDF_CHECKS)
#define FDF_BOUND(i, n)                                                                            \
    do {                                                                                           \
        if ((unsigned long long)(i) >= (unsigned long long)(n)) atomicMax(&::fdf::g_check_failure_line, __LINE__); \
    } while (0)
#else
#define FDF_BOUND(i, n) \
    do {                \
    } while (0)
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

FDF_HD int popc32(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}

FDF_HD int highest_set_bit(uint32_t m) {  // m != 0                              FLO.U32
#if defined(__CUDA_ARCH__)
    int r;
    asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(m));
    return r;
#else
    return 31 - __builtin_clz(m);
#endif
}

// a * b + c on the multiply-add pipe.  ptxas is free to turn small cases into adds / shifts; the point of
// spelling it this way is that it never needs the logic pipe.
FDF_HD uint32_t mad32(uint32_t a, uint32_t b, uint32_t c) { return a * b + c; }

// ---- 2 x 16-bit lane helpers (DPX: VIMNMX3.U16x2, VIMNMX.U16x2, VIADDMNMX.S16x2.RELU) ----------
FDF_HD uint32_t swap16(uint32_t v) { return byte_perm(v, 0u, 0x1032u); }

FDF_HD uint32_t min_u16x2(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __vminu2(a, b);
#else
    return min(a & 0xffffu, b & 0xffffu) | (min(a >> 16, b >> 16) << 16);
#endif
}
FDF_HD uint32_t max_u16x2(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __vmaxu2(a, b);
#else
    return max(a & 0xffffu, b & 0xffffu) | (max(a >> 16, b >> 16) << 16);
#endif
}
FDF_HD uint32_t min3_u16x2(uint32_t a, uint32_t b, uint32_t c) {
#if defined(__CUDA_ARCH__)
    return __vimin3_u16x2(a, b, c);
#else
    return min_u16x2(a, min_u16x2(b, c));
#endif
}
FDF_HD uint32_t max3_u16x2(uint32_t a, uint32_t b, uint32_t c) {
#if defined(__CUDA_ARCH__)
    return __vimax3_u16x2(a, b, c);
#else
    return max_u16x2(a, max_u16x2(b, c));
#endif
}
// per 16-bit lane: max(a + b, 0) with signed wrap-around lanes
FDF_HD uint32_t addrelu_s16x2(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __viaddmax_s16x2_relu(a, b, 0u);
#else
    uint32_t r = 0;
    for (int i = 0; i < 2; i++) {
        const int16_t s = (int16_t)(uint16_t)(((a >> (16 * i)) + (b >> (16 * i))) & 0xffffu);
        r |= (uint32_t)(uint16_t)(s > 0 ? s : 0) << (16 * i);
    }
    return r;
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
// Threshold test of a word m of four byte values:  f = (m + kbias) | m, flag = bit 7 of each byte, with
// kbias = (0x7f - t) * 0x01010101 for t < 128 and 0 for t >= 128.  Per byte (t < 128): if m <= 0x7f the
// sum cannot carry out and has bit 7 set iff m + 0x7f - t >= 0x80 iff m > t; if m >= 0x80 (> t) the OR
// with m sets bit 7.  A carry out of a byte >= 0x80 adds 1 to the next byte, which can only turn a flag
// ON: the test stays a necessary condition and no byte mask is needed.  For t >= 128 the flag is
// m >= 128, implied by m > t.  All other bits of f are garbage.
FDF_HD uint32_t filter_kbias(uint32_t t) { return t < 128u ? (0x7fu - t) * 0x01010101u : 0u; }

FDF_HD uint32_t exceeds4(uint32_t m, uint32_t kbias) { return mad32(m, 1u, kbias) | m; }

// the 16 pixels x .. x+15 of a tile row, as four little-endian words
struct Px16 {
    uint32_t w[4];
};

FDF_HD Px16 load16(const uint8_t *p) {  // p 16-byte aligned                      LDS.128
    const uint4 v = *reinterpret_cast<const uint4 *>(p);
    Px16 r;
    r.w[0] = v.x;
    r.w[1] = v.y;
    r.w[2] = v.z;
    r.w[3] = v.w;
    return r;
}

// Stage 2 (16-pixel groups that passed stage 1): both pairs.  cl / cr are the words left / right of the
// centre row's 16 pixels; `valid` is the group's mask of pixels that may be a centre at all (image border, chunk
// halo), in the layout of the result.  Returns the candidate mask of the group: centre 4k + b  <->  bit 8b + 7 - k.
// The east/west differences are computed once per pixel: H(x) = |p(x+3) - p(x)| is the east difference
// of centre x and the west difference of centre x + 3.
FDF_HD uint32_t candidate_mask16(const Px16 &c, const Px16 &n, const Px16 &s, uint32_t cl, uint32_t cr,
                                 uint32_t valid, uint32_t kbias) {
    uint32_t h[5];  // h[k + 1] = H of word k, h[0] = H of the word left of the group
    h[0] = absdiff4(byte_perm(cl, c.w[0], 0x6543u), cl);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint32_t east = byte_perm(c.w[k], k < 3 ? c.w[k + 1] : cr, 0x6543u);  // pixels x+3 .. x+6
        h[k + 1] = absdiff4(east, c.w[k]);
    }
    uint32_t r[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint32_t a = absdiff4(n.w[k], c.w[k]) | absdiff4(s.w[k], c.w[k]);
        const uint32_t b = h[k + 1] | byte_perm(h[k], h[k + 1], 0x4321u);  // H(x) | H(x-3)
        r[k] = exceeds4(a, kbias) & exceeds4(b, kbias) & 0x80808080u;  // only bit 7 of every byte is meaningful
    }
    return (mad32(r[0], 1u, r[1] >> 1) + mad32(r[2] >> 2, 1u, r[3] >> 3)) & valid;
}

// candidate-mask bit -> pixel index inside the group, and back
FDF_HD int mask_bit_to_px(int p) { return 4 * (7 - (p & 7)) + (p >> 3); }
FDF_HD int px_to_mask_bit(int px) { return 8 * (px & 3) + 7 - (px >> 2); }

// ---- exact test + scores on 16 "dual" words (the form the detection kernel uses) ---------------------------
//
// One word per ring pixel:  lane 0 = 256 + (p_i - c),  lane 1 = 256 - (p_i - c)   (both in [1, 511]).
// A single multiply-add builds it straight from the loaded byte:  p * (1 - 2^16) + ((256 + c) << 16 | (256 - c)),
// so there is no separate packing step and no lane swap anywhere below: the cyclic windows are plain index
// arithmetic mod 16.
//
// Segment test and MaxThreshold score in one go.  With W_k the cyclic window of n ring positions starting at k:
//   lane 0 of max_k min_{W_k} word  =  256 + max_k min_{W_k} (p - c)   > 256 + t  <=>  a brighter arc of >= n
//   lane 1                          =  256 + max_k min_{W_k} (c - p)   > 256 + t  <=>  a darker arc of >= n
// (fast_simd.rs:218-296 asks for exactly that: some window whose pixels are all > c + t, or all < c - t.)
// Any two windows of >= 9 of 16 positions overlap, so when one lane exceeds 256 + t the other one is below 256:
// best = max(lane 0, lane 1), and  keypoint <=> best > 256 + t.
// MaxThreshold score (opencv_compat.rs:172-209): with d_i = c - p_i,  eh = max_k min_{W_k} d,  el = min_k max_{W_k} d,
// score = min(|eh|, |el|).  Overlapping windows give eh <= el.  For a keypoint with a brighter arc el <= -(t+1) < 0, so
// eh <= el < 0 and the score is -el = max_k min_{W_k} (p - c) = lane 0 - 256; for a darker arc eh >= t+1 > 0, so
// el >= eh > 0 and the score is eh = lane 1 - 256.  Hence best - 256 IS the score of a keypoint (the only pixels
// that are ever scored, fast_simd.rs:276-279).
struct RingDual {
    uint32_t w[16];
};

FDF_HD uint32_t dual_bias(int c) { return ((uint32_t)(256 + c) << 16) | (uint32_t)(256 - c); }
FDF_HD uint32_t dual_word(uint32_t p, uint32_t bias) { return mad32(p, 0xffff0001u, bias); }

template <int K>  // K = n - 9
FDF_HD uint32_t best_window_k(const uint32_t u[16]) {
    uint32_t v[16];
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = (K == 0) ? u[i] : min_u16x2(u[i], u[(i + K) & 15]);
    const uint32_t a0 = max3_u16x2(v[0], v[1], v[2]), a1 = max3_u16x2(v[3], v[4], v[5]);
    const uint32_t a2 = max3_u16x2(v[6], v[7], v[8]), a3 = max3_u16x2(v[9], v[10], v[11]);
    const uint32_t a4 = max3_u16x2(v[12], v[13], v[14]);
    return max_u16x2(max3_u16x2(a0, a1, a2), max3_u16x2(a3, a4, v[15]));
}

// per lane: 256 + max over the 16 cyclic windows of n positions of the window's minimum
// NFIX: n as a compile-time constant (9..16), or 0 = read it at run time.  The detection kernel instantiates its
// per-candidate loop for n = 9 (the reference's and OpenCV's default) and once more for "any n": the run-time form costs
// an indirect branch per candidate.
template <int NFIX>
FDF_HD uint32_t best_window(const RingDual &r, int n) {
    uint32_t t3[16], u[16];
#pragma unroll
    for (int i = 0; i < 16; i++) t3[i] = min3_u16x2(r.w[i], r.w[(i + 1) & 15], r.w[(i + 2) & 15]);
#pragma unroll
    for (int i = 0; i < 16; i++) u[i] = min3_u16x2(t3[i], t3[(i + 3) & 15], t3[(i + 6) & 15]);  // 9-windows
    if (NFIX >= 9) return best_window_k<(NFIX >= 9 ? NFIX - 9 : 0)>(u);
    switch (n) {
        case 9: return best_window_k<0>(u);
        case 10: return best_window_k<1>(u);
        case 11: return best_window_k<2>(u);
        case 12: return best_window_k<3>(u);
        case 13: return best_window_k<4>(u);
        case 14: return best_window_k<5>(u);
        case 15: return best_window_k<6>(u);
        default: return best_window_k<7>(u);
    }
}

// max of the two lanes of best_window: > 256 + t  <=>  keypoint;  minus 256 = MaxThreshold score of a keypoint
FDF_HD uint32_t best_of_lanes(uint32_t m) { return max(m & 0xffffu, m >> 16); }

// SumAbsolute on dual words: lane 0 accumulates max(p - c - t, 0), lane 1 max(c - p - t, 0)  (opencv_compat.rs:278-299)
FDF_HD uint32_t score_sum_abs_dual(const RingDual &r, int t) {
    const uint32_t k = (uint32_t)((0x10000 - (256 + t)) & 0xffff) * 0x00010001u;  // -(256 + t) per lane
    uint32_t s = 0u;
#pragma unroll
    for (int i = 0; i < 16; i++) s += addrelu_s16x2(r.w[i], k);  // lane sums <= 16 * 255
    return max(s & 0xffffu, s >> 16);
}

}  // namespace fdf
