// second_impl.h -- an independent second implementation of the per-candidate arithmetic (the detection kernel's
// first generation: two ring pixels per word, separate brighter / darker masks, rotate-AND arc test, separate score
// passes).  TEST CODE: strip_emulator.cpp::fdf_core_check compares it against the oracle AND against the dual-word
// forms the kernel uses (csrc/fdf_core.cuh).  It is not part of the library.
#pragma once
#include "../../feature_detector_fast_b200/csrc/fdf_core.cuh"

namespace fdf {

// Stage 1 (every scored pixel): north/south pair only.  Returns non-zero iff at least one of the 16 centres
// has a north or south ring pixel that differs from it by more than t.
FDF_HD uint32_t vertical_any(const Px16 &c, const Px16 &n, const Px16 &s, uint32_t kbias) {
    uint32_t o = 0u;
#pragma unroll
    for (int k = 0; k < 4; k++) o |= exceeds4(absdiff4(n.w[k], c.w[k]) | absdiff4(s.w[k], c.w[k]), kbias);
    return o & 0x80808080u;
}

// ---- exact segment test on 2 x 16-bit lanes ----------------------------------------------------------
// (ring_masks / has_arc / score_max_threshold / score_sum_abs below are the first-generation forms: the kernel now uses
// the dual-word forms at the end of this file; these stay as an independent second implementation that
// tests/host/strip_emulator.cpp::fdf_core_check compares against the oracle AND against the dual-word forms.)
// The 16 ring pixels are held as 8 words  P[i] = ring[i] | ring[i + 8] << 16  (opposite pixels share a word).
struct Ring2 {
    uint32_t p[8];
};

struct RingMasks {
    uint32_t bright;  // bit i set <=> ring[i] > c + t   (fast_simd.rs:224 is_above)
    uint32_t dark;    // bit i set <=> ring[i] < c - t   (fast_simd.rs:225 is_below)
};

// bright: lane value p + 255 - hi with hi = min(c + t, 255) lies in [0, 510] and has bit 8 set iff p > hi;
// dark:   lane value 255 + lo - p with lo = max(c - t, 0)  lies in [0, 510] and has bit 8 set iff p < lo.
// Neither can borrow across lanes.  The flags are collected Horner style (acc = 2 * acc + flag), so ring i
// ends up in bit 8 + i of its lane; one PRMT then gathers the two lane bytes into a 16-bit mask.
FDF_HD RingMasks ring_masks(int c, const Ring2 &r, int t) {
    const uint32_t kb = (uint32_t)(255 - min(c + t, 255)) * 0x00010001u;
    const uint32_t kd = (uint32_t)(255 + max(c - t, 0)) * 0x00010001u;
    uint32_t ab = 0u, ad = 0u;
#pragma unroll
    for (int i = 7; i >= 0; i--) {
        ab = mad32(ab, 2u, (r.p[i] + kb) & 0x01000100u);
        ad = mad32(ad, 2u, (kd - r.p[i]) & 0x01000100u);
    }
    RingMasks m;
    m.bright = byte_perm(ab, 0u, 0x4431u);
    m.dark = byte_perm(ad, 0u, 0x4431u);
    return m;
}

// exists a cyclic run of >= n set bits in the 16-bit ring mask (9 <= n <= 16)
FDF_HD bool has_arc(uint32_t m16, int n) {
    uint32_t r = mad32(m16, 0x00010001u, 0u);
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
// e_i = +-(c - p_i) is exact.  Here e is biased by 256 (lanes in [1, 511], unsigned), two ring positions
// 8 apart per word:  E[i] = (e_i, e_{i+8}),  E[i + 8] = swap16(E[i]).  Window minima: 3-window, then
// 9-window = min3 of three 3-windows, then n-window = min(9-window at k, 9-window at k + n - 9).
template <int K>
FDF_HD uint32_t max_of_extended(const uint32_t u[16]) {  // u[i + 8] = swap16(u[i]); only u[0 .. 7 + K] are read
    uint32_t v[8];
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = (K == 0) ? u[i] : min_u16x2(u[i], u[i + K]);
    const uint32_t a = max3_u16x2(v[0], v[1], v[2]), b = max3_u16x2(v[3], v[4], v[5]);
    const uint32_t m = max3_u16x2(a, b, max_u16x2(v[6], v[7]));
    return max(m & 0xffffu, m >> 16);
}

// pixel_is_brighter: the arc found by the segment test is a "bright" arc (ring pixels > c + t)
FDF_HD uint32_t score_max_threshold(int c, const Ring2 &r, int n, bool pixel_is_brighter) {
    const uint32_t sgn = pixel_is_brighter ? 1u : 0xffffffffu;
    const uint32_t k = pixel_is_brighter ? (uint32_t)(256 - c) * 0x00010001u : (uint32_t)(256 + c) * 0x00010001u;
    uint32_t e[10], t3[14], u[16];
#pragma unroll
    for (int i = 0; i < 8; i++) e[i] = mad32(r.p[i], sgn, k);  // 256 +- (p - c) per lane, no cross-lane borrow
    e[8] = swap16(e[0]);
    e[9] = swap16(e[1]);
#pragma unroll
    for (int i = 0; i < 8; i++) t3[i] = min3_u16x2(e[i], e[i + 1], e[i + 2]);
#pragma unroll
    for (int i = 0; i < 6; i++) t3[i + 8] = swap16(t3[i]);
#pragma unroll
    for (int i = 0; i < 8; i++) u[i] = min3_u16x2(t3[i], t3[i + 3], t3[i + 6]);
    uint32_t m;
    if (n == 9) {
        m = max_of_extended<0>(u);
    } else {
#pragma unroll
        for (int i = 0; i < 7; i++) u[i + 8] = swap16(u[i]);
        u[15] = 0u;
        switch (n) {
            case 10: m = max_of_extended<1>(u); break;
            case 11: m = max_of_extended<2>(u); break;
            case 12: m = max_of_extended<3>(u); break;
            case 13: m = max_of_extended<4>(u); break;
            case 14: m = max_of_extended<5>(u); break;
            case 15: m = max_of_extended<6>(u); break;
            default: m = max_of_extended<7>(u); break;
        }
    }
    return m - 256u;
}

// SumAbsolute (opencv_compat.rs:278-299): max( sum_{p > c+t} (p-c-t), sum_{p < c-t} (c-p-t) ) over
// ALL 16 ring pixels.  p - c - t > 0 <=> p > c + t, so each term is a relu; clamping c + t to 255 and
// c - t to 0 changes nothing (every term is then <= 0).  Lane sums stay below 8 * 255.
FDF_HD uint32_t score_sum_abs(int c, const Ring2 &r, int t) {
    const uint32_t nhi = (uint32_t)((0x10000 - min(c + t, 255)) & 0xffff) * 0x00010001u;  // -hi per lane
    const uint32_t lo1 = (uint32_t)(max(c - t, 0) + 1) * 0x00010001u;                      // lo + 1 per lane
    uint32_t sb = 0u, sd = 0u;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        sb += addrelu_s16x2(r.p[i], nhi);   // max(p - hi, 0)
        sd += addrelu_s16x2(~r.p[i], lo1);  // (-p - 1) + (lo + 1) = lo - p
    }
    sb = (sb & 0xffffu) + (sb >> 16);
    sd = (sd & 0xffffu) + (sd >> 16);
    return max(sb, sd);
}


}  // namespace fdf
