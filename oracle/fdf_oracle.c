/*
 * fdf_oracle.c -- scalar CPU restatement of the reference's FAST-n specification.
 *
 * TEST INFRASTRUCTURE ONLY (see fdf_oracle.h).  Plain C, no SIMD, written for clarity: every
 * function names the reference lines it follows.  All arithmetic is integer (u8 pixels, i16/i32
 * differences, u16 scores), exactly as in src/opencv_compat.rs.
 */
#include "fdf_oracle.h"

#include <stdlib.h>
#include <string.h>

/* opencv_compat.rs:42-61 (same table as fast_simd.rs:79-98): (dx, dy), y grows downwards. */
static const int CIRCLE[16][2] = {
    {0, -3}, {1, -3}, {2, -2}, {3, -1}, {3, 0},  {3, 1},   {2, 2},   {1, 3},
    {0, 3},  {-1, 3}, {-2, 2}, {-3, 1}, {-3, 0}, {-3, -1}, {-2, -2}, {-1, -3},
};

void fdf_oracle_circle(int32_t out_dxdy[32]) {
    for (int i = 0; i < 16; i++) {
        out_dxdy[2 * i] = CIRCLE[i][0];
        out_dxdy[2 * i + 1] = CIRCLE[i][1];
    }
}

/* opencv_compat.rs:140-165: for each start s, walk the ring while flags are set; a keypoint
 * needs some start with a run of at least n.  (iter().cycle().skip(s).take(len).take_while()) */
int fdf_oracle_consecutive(const uint8_t *flags, int len, int n) {
    for (int s = 0; s < len; s++) {
        int run = 0;
        while (run < len && flags[(s + run) % len]) run++;
        if (run >= n) return 1;
    }
    return 0;
}

static void gather_circle(const uint8_t *img, uint32_t pitch, uint32_t x, uint32_t y,
                          uint8_t circle[16]) {
    for (int i = 0; i < 16; i++) {
        circle[i] = img[(size_t)((int64_t)y + CIRCLE[i][1]) * pitch + (size_t)((int64_t)x + CIRCLE[i][0])];
    }
}

/* opencv_compat.rs:95-166.  delta = centre - pixel (i16); neg <=> delta < 0 && |delta| > t,
 * pos <=> delta > 0 && |delta| > t; keypoint <=> a cyclic run >= n in neg or in pos. */
int fdf_oracle_is_keypoint(const uint8_t *img, uint32_t pitch, uint32_t x, uint32_t y, uint8_t t,
                           uint8_t n) {
    uint8_t circle[16];
    uint8_t neg[16], pos[16];
    gather_circle(img, pitch, x, y, circle);
    int base = img[(size_t)y * pitch + x];
    for (int i = 0; i < 16; i++) {
        int d = base - (int)circle[i];
        int a = d < 0 ? -d : d;
        neg[i] = (uint8_t)(d < 0 && a > (int)t);
        pos[i] = (uint8_t)(d > 0 && a > (int)t);
    }
    return fdf_oracle_consecutive(neg, 16, n) || fdf_oracle_consecutive(pos, 16, n);
}

/* opencv_compat.rs:172-209.  difference[i] = centre - circle[i % 16] for i < 32;
 * extreme_highest = max_k min(difference[k..k+n]); extreme_lowest = min_k max(difference[k..k+n]);
 * score = min(|extreme_highest|, |extreme_lowest|). */
uint16_t fdf_oracle_score_max_threshold_px(uint8_t centre, const uint8_t circle[16], uint8_t n) {
    int diff[32];
    for (int i = 0; i < 32; i++) diff[i] = (int)centre - (int)circle[i % 16];
    int extreme_highest = -32768;
    int extreme_lowest = 32767;
    for (int k = 0; k < 16; k++) {
        int mn = diff[k], mx = diff[k];
        for (int j = 1; j < (int)n; j++) {
            if (diff[k + j] < mn) mn = diff[k + j];
            if (diff[k + j] > mx) mx = diff[k + j];
        }
        if (mn > extreme_highest) extreme_highest = mn;
        if (mx < extreme_lowest) extreme_lowest = mx;
    }
    int a = extreme_highest < 0 ? -extreme_highest : extreme_highest;
    int b = extreme_lowest < 0 ? -extreme_lowest : extreme_lowest;
    return (uint16_t)(a < b ? a : b);
}

uint16_t fdf_oracle_score_max_threshold(const uint8_t *img, uint32_t pitch, uint32_t x, uint32_t y,
                                        uint8_t n) {
    uint8_t circle[16];
    gather_circle(img, pitch, x, y, circle);
    return fdf_oracle_score_max_threshold_px(img[(size_t)y * pitch + x], circle, n);
}

/* opencv_compat.rs:278-299.  Sums run over ALL 16 circle pixels beyond the threshold, not only
 * the arc.  ("light" = centre brighter than pixel, "dark" = pixel brighter; names as upstream.) */
uint16_t fdf_oracle_score_sum_abs_px(uint8_t centre, const uint8_t circle[16], uint8_t t) {
    unsigned sum_dark = 0, sum_light = 0;
    for (int i = 0; i < 16; i++) {
        int d = (int)centre - (int)circle[i];
        int a = d < 0 ? -d : d;
        if (d > 0 && a > (int)t) sum_light += (unsigned)(centre - circle[i]) - t;
        if (d < 0 && a > (int)t) sum_dark += (unsigned)(circle[i] - centre) - t;
    }
    return (uint16_t)(sum_dark > sum_light ? sum_dark : sum_light);
}

uint16_t fdf_oracle_score_sum_abs(const uint8_t *img, uint32_t pitch, uint32_t x, uint32_t y,
                                  uint8_t t) {
    uint8_t circle[16];
    gather_circle(img, pitch, x, y, circle);
    return fdf_oracle_score_sum_abs_px(img[(size_t)y * pitch + x], circle, t);
}

/*
 * opencv_compat.rs:302-306 -> detect (:79-169) then non_max_supression (:212-262).
 *
 * The reference's NMS searches the keypoint Vec for each of the 8 neighbours (O(K^2)); here the
 * same predicate is evaluated through a dense "is keypoint" score plane (0 = no keypoint), which
 * is what fast_simd.rs:588-616 does with its three rolling rows.  The predicate itself is kept
 * verbatim: rows 3 and h-4 are dropped (:238-240), only neighbours that are keypoints count
 * (:249), and a keypoint is removed when current <= other (:254).
 */
int64_t fdf_oracle_detect(const uint8_t *img, uint32_t w, uint32_t h, uint32_t pitch, uint8_t t,
                          uint8_t n, uint8_t nms, fdf_oracle_point *out, size_t cap,
                          uint16_t *scores_out) {
    if (n < 9 || n > 16) return -1; /* fast_simd.rs:302-305 (assert), :797-801 (OOB panic) */
    if (nms > FDF_ORACLE_NMS_SUM_ABSOLUTE) return -2;
    if (w < 7 || h < 7) return 0; /* SURVEY S15 */

    int64_t count = 0;
    if (nms == FDF_ORACLE_NMS_OFF) {
        for (uint32_t y = 3; y < h - 3; y++) {
            for (uint32_t x = 3; x < w - 3; x++) {
                if (!fdf_oracle_is_keypoint(img, pitch, x, y, t, n)) continue;
                if ((size_t)count < cap) {
                    out[count].x = x;
                    out[count].y = y;
                    if (scores_out) scores_out[count] = 0;
                }
                count++;
            }
        }
        return count;
    }

    /* score plane: 0 <=> not a keypoint (every keypoint scores >= 1) */
    uint16_t *plane = (uint16_t *)calloc((size_t)w * h, sizeof(uint16_t));
    if (!plane) return -3;
    for (uint32_t y = 3; y < h - 3; y++) {
        for (uint32_t x = 3; x < w - 3; x++) {
            if (!fdf_oracle_is_keypoint(img, pitch, x, y, t, n)) continue;
            uint16_t s = (nms == FDF_ORACLE_NMS_MAX_THRESHOLD)
                             ? fdf_oracle_score_max_threshold(img, pitch, x, y, n)
                             : fdf_oracle_score_sum_abs(img, pitch, x, y, t);
            plane[(size_t)y * w + x] = s;
        }
    }
    for (uint32_t y = 3; y < h - 3; y++) {
        if (y == 3 || y == h - 4) continue; /* opencv_compat.rs:238-240 */
        for (uint32_t x = 3; x < w - 3; x++) {
            uint16_t cur = plane[(size_t)y * w + x];
            if (cur == 0) continue;
            int keep = 1;
            for (int dx = -1; dx <= 1 && keep; dx++) {
                for (int dy = -1; dy <= 1; dy++) {
                    if (dx == 0 && dy == 0) continue;
                    uint16_t other = plane[(size_t)((int64_t)y + dy) * w + (size_t)((int64_t)x + dx)];
                    if (other == 0) continue; /* neighbour is not a keypoint: :249 */
                    if (cur <= other) {        /* :254 */
                        keep = 0;
                        break;
                    }
                }
            }
            if (!keep) continue;
            if ((size_t)count < cap) {
                out[count].x = x;
                out[count].y = y;
                if (scores_out) scores_out[count] = cur;
            }
            count++;
        }
    }
    free(plane);
    return count;
}

/* ---- SipHash-1-3, key (0,0): Rust's std DefaultHasher, as used by tests/compare.rs:5-20 ---- */
#define ROTL64(x, b) (((x) << (b)) | ((x) >> (64 - (b))))
#define SIPROUND            \
    do {                    \
        v0 += v1;           \
        v1 = ROTL64(v1, 13); \
        v1 ^= v0;           \
        v0 = ROTL64(v0, 32); \
        v2 += v3;           \
        v3 = ROTL64(v3, 16); \
        v3 ^= v2;           \
        v0 += v3;           \
        v3 = ROTL64(v3, 21); \
        v3 ^= v0;           \
        v2 += v1;           \
        v1 = ROTL64(v1, 17); \
        v1 ^= v2;           \
        v2 = ROTL64(v2, 32); \
    } while (0)

uint64_t fdf_oracle_siphash13(const uint8_t *data, size_t len) {
    uint64_t v0 = 0x736f6d6570736575ULL, v1 = 0x646f72616e646f6dULL;
    uint64_t v2 = 0x6c7967656e657261ULL, v3 = 0x7465646279746573ULL;
    size_t blocks = len / 8;
    for (size_t i = 0; i < blocks; i++) {
        uint64_t m = 0;
        for (int b = 0; b < 8; b++) m |= (uint64_t)data[i * 8 + b] << (8 * b);
        v3 ^= m;
        SIPROUND;
        v0 ^= m;
    }
    uint64_t tail = (uint64_t)(len & 0xff) << 56;
    for (size_t b = 0; b < (len & 7); b++) tail |= (uint64_t)data[blocks * 8 + b] << (8 * b);
    v3 ^= tail;
    SIPROUND;
    v0 ^= tail;
    v2 ^= 0xff;
    SIPROUND;
    SIPROUND;
    SIPROUND;
    return v0 ^ v1 ^ v2 ^ v3;
}

/* <[Point] as Hash>::hash: length prefix (usize, 8 bytes LE) then each derive(Hash) field. */
uint64_t fdf_oracle_hash_points(const fdf_oracle_point *pts, size_t n) {
    size_t len = 8 + n * 8;
    uint8_t *buf = (uint8_t *)malloc(len ? len : 1);
    if (!buf) return 0;
    uint64_t n64 = (uint64_t)n;
    for (int b = 0; b < 8; b++) buf[b] = (uint8_t)(n64 >> (8 * b));
    for (size_t i = 0; i < n; i++) {
        for (int b = 0; b < 4; b++) {
            buf[8 + i * 8 + b] = (uint8_t)(pts[i].x >> (8 * b));
            buf[8 + i * 8 + 4 + b] = (uint8_t)(pts[i].y >> (8 * b));
        }
    }
    uint64_t r = fdf_oracle_siphash13(buf, len);
    free(buf);
    return r;
}

/* ---- synthetic frames: a pure function of (seed, frame, x, y); mirrored in fdf_synth.cuh ---- */
static uint64_t mix64(uint64_t z) { /* splitmix64 finaliser */
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

static uint64_t hash4(uint64_t key, uint32_t a, uint32_t b, uint32_t c) {
    return mix64(key ^ ((uint64_t)a * 0xD6E8FEB86659FD93ULL) ^ ((uint64_t)b * 0xA0761D6478BD642FULL) ^
                 ((uint64_t)c * 0xE7037ED1A0B428DBULL));
}

/* One rectangle layer: cells of `cell` px (shifted by a per-frame offset); a cell is active with
 * probability prob/256 and then paints an inset rectangle with its own grey level. */
static int layer_level(uint64_t key, uint32_t layer, uint32_t x, uint32_t y, uint32_t cell,
                       uint32_t prob, uint32_t max_inset, int *level) {
    uint64_t ho = hash4(key, layer, 0xFFFFFFFFu, 0xFFFFFFFEu);
    uint32_t ox = (uint32_t)(ho % cell), oy = (uint32_t)((ho >> 20) % cell);
    uint32_t cx = (x + ox) / cell, cy = (y + oy) / cell;
    uint32_t px = (x + ox) % cell, py = (y + oy) % cell;
    uint64_t hc = hash4(key, layer, cx, cy);
    if ((uint32_t)(hc & 0xff) >= prob) return 0;
    uint32_t l = (uint32_t)((hc >> 8) % (max_inset + 1));
    uint32_t r = (uint32_t)((hc >> 16) % (max_inset + 1));
    uint32_t tp = (uint32_t)((hc >> 24) % (max_inset + 1));
    uint32_t bt = (uint32_t)((hc >> 32) % (max_inset + 1));
    if (px < l || px >= cell - r || py < tp || py >= cell - bt) return 0;
    *level = 16 + (int)((hc >> 40) % 224);
    return 1;
}

void fdf_oracle_synth_frame(uint8_t *out, uint32_t w, uint32_t h, uint32_t pitch, uint64_t seed,
                            uint32_t frame, uint32_t kind, uint32_t amp) {
    uint64_t key = mix64(seed ^ ((uint64_t)frame * 0x8CB92BA72F3D8DD7ULL));
    for (uint32_t y = 0; y < h; y++) {
        for (uint32_t x = 0; x < w; x++) {
            uint64_t hp = hash4(key, 7u, x, y);
            int v;
            if (kind == 1) {
                v = (int)(hp & 0xff);
            } else {
                int level = 32 + (int)((hash4(key, 0u, x / 96, y / 96) >> 8) % 192);
                int lv;
                if (layer_level(key, 1u, x, y, 40u, 72u, 12u, &lv)) level = lv;
                if (layer_level(key, 2u, x, y, 16u, 20u, 5u, &lv)) level = lv;
                int noise = (int)(hp % (2 * amp + 1)) - (int)amp;
                v = level + noise;
                if (v < 0) v = 0;
                if (v > 255) v = 255;
            }
            out[(size_t)y * pitch + x] = (uint8_t)v;
        }
    }
}
