// fdf_synth.cuh -- counter-based synthetic grey frames for tests and benchmarks.
//
// Every pixel is a pure integer function of (seed, frame, x, y, kind, amp), so the GPU generator
// below and the CPU oracle's independent copy (oracle/fdf_oracle.c: fdf_oracle_synth_frame) produce
// identical bytes; tests/test_gpu_parity.py checks that before any frame is used for parity.
//   kind 0 "scene": 96-px background blocks + two layers of inset random rectangles (40-px and
//                   16-px cells) + uniform noise in [-amp, +amp].  Tuned so that a 1080p frame gives
//                   about as many keypoints at t=16, n=9, NMS off as the reference's published
//                   1080p frame (README.md:58-59: 23184).
//   kind 1 "noise": uniform random bytes (stress case, ~28 % keypoints).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define FDF_SYNTH_HD __host__ __device__ __forceinline__
#else
#define FDF_SYNTH_HD inline
#endif

namespace fdf {

FDF_SYNTH_HD uint64_t synth_mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

FDF_SYNTH_HD uint64_t synth_hash(uint64_t key, uint32_t a, uint32_t b, uint32_t c) {
    return synth_mix64(key ^ ((uint64_t)a * 0xD6E8FEB86659FD93ULL) ^ ((uint64_t)b * 0xA0761D6478BD642FULL) ^
                       ((uint64_t)c * 0xE7037ED1A0B428DBULL));
}

FDF_SYNTH_HD uint64_t synth_frame_key(uint64_t seed, uint32_t frame) {
    return synth_mix64(seed ^ ((uint64_t)frame * 0x8CB92BA72F3D8DD7ULL));
}

FDF_SYNTH_HD bool synth_layer(uint64_t key, uint32_t layer, uint32_t x, uint32_t y, uint32_t cell, uint32_t prob,
                              uint32_t max_inset, int *level) {
    const uint64_t ho = synth_hash(key, layer, 0xFFFFFFFFu, 0xFFFFFFFEu);
    const uint32_t ox = (uint32_t)(ho % cell), oy = (uint32_t)((ho >> 20) % cell);
    const uint32_t cx = (x + ox) / cell, cy = (y + oy) / cell;
    const uint32_t px = (x + ox) % cell, py = (y + oy) % cell;
    const uint64_t hc = synth_hash(key, layer, cx, cy);
    if ((uint32_t)(hc & 0xff) >= prob) return false;
    const uint32_t l = (uint32_t)((hc >> 8) % (max_inset + 1));
    const uint32_t r = (uint32_t)((hc >> 16) % (max_inset + 1));
    const uint32_t tp = (uint32_t)((hc >> 24) % (max_inset + 1));
    const uint32_t bt = (uint32_t)((hc >> 32) % (max_inset + 1));
    if (px < l || px >= cell - r || py < tp || py >= cell - bt) return false;
    *level = 16 + (int)((hc >> 40) % 224);
    return true;
}

FDF_SYNTH_HD uint8_t synth_pixel(uint64_t key, uint32_t x, uint32_t y, uint32_t kind, uint32_t amp) {
    const uint64_t hp = synth_hash(key, 7u, x, y);
    if (kind == 1u) return (uint8_t)(hp & 0xff);
    int level = 32 + (int)((synth_hash(key, 0u, x / 96u, y / 96u) >> 8) % 192);
    int lv;
    if (synth_layer(key, 1u, x, y, 40u, 72u, 12u, &lv)) level = lv;
    if (synth_layer(key, 2u, x, y, 16u, 20u, 5u, &lv)) level = lv;
    int v = level + (int)(hp % (2u * amp + 1u)) - (int)amp;
    v = v < 0 ? 0 : (v > 255 ? 255 : v);
    return (uint8_t)v;
}

}  // namespace fdf
