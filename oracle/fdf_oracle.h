/*
 * fdf_oracle.h -- CPU oracle for the FAST-n detection path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it,
 * and only as the checker / CPU baseline.  The product path (libfdf_cuda.so) never links or
 * calls anything in here.
 *
 * What it restates (citations into the reference checkout, /root/reference):
 *   - scalar specification         src/opencv_compat.rs:42-306
 *   - plane / ordering behaviour   src/fast_simd.rs:588-616
 *   - AVX2 hot path (the "port")   src/fast_simd.rs:115-297, 301-620, 623-824   (fdf_avx2_port.cpp)
 *
 * Parity status: PINNED against the reference's own known-answer vector
 * (fast_simd.rs:919-937, 965-1021 -> score 20), the shipped golden renders in media/
 * (309 / 131 keypoints, tests/golden/), the test_consecutive vectors (opencv_compat.rs:327-345)
 * and cv2 4.13 FAST(TYPE_9_16).  The reference itself cannot be compiled here (no Rust toolchain).
 */
#ifndef FDF_ORACLE_H
#define FDF_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    uint32_t x;
    uint32_t y;
} fdf_oracle_point; /* lib.rs:15-20 */

enum {
    FDF_ORACLE_NMS_OFF = 0,           /* fast_simd.rs:74 */
    FDF_ORACLE_NMS_MAX_THRESHOLD = 1, /* fast_simd.rs:75 */
    FDF_ORACLE_NMS_SUM_ABSOLUTE = 2   /* fast_simd.rs:76 */
};

/* The 16 circle offsets (dx, dy), index 0 = north, clockwise.  opencv_compat.rs:42-61 */
void fdf_oracle_circle(int32_t out_dxdy[32]);

/* Cyclic "exists a run of >= n set flags" on a ring of len flags.  opencv_compat.rs:140-165, 312-325 */
int fdf_oracle_consecutive(const uint8_t *flags, int len, int n);

/* Segment test for one centre.  opencv_compat.rs:95-166 */
int fdf_oracle_is_keypoint(const uint8_t *img, uint32_t pitch, uint32_t x, uint32_t y, uint8_t t,
                           uint8_t n);

/* MaxThreshold score from a centre and its 16 circle pixels.  opencv_compat.rs:172-209 */
uint16_t fdf_oracle_score_max_threshold_px(uint8_t centre, const uint8_t circle[16], uint8_t n);
uint16_t fdf_oracle_score_max_threshold(const uint8_t *img, uint32_t pitch, uint32_t x, uint32_t y,
                                        uint8_t n);

/* SumAbsolute score.  opencv_compat.rs:278-299 */
uint16_t fdf_oracle_score_sum_abs_px(uint8_t centre, const uint8_t circle[16], uint8_t t);
uint16_t fdf_oracle_score_sum_abs(const uint8_t *img, uint32_t pitch, uint32_t x, uint32_t y,
                                  uint8_t t);

/*
 * Whole-image detector: opencv_compat.rs:302-306 (detect + non_max_supression).
 * Returns the number of keypoints found (which may exceed cap; only the first cap are written),
 * or a negative value for an invalid count (the reference panics: fast_simd.rs:302-305, :797-801).
 * Images with w < 7 or h < 7 give 0.
 * scores_out (optional, may be NULL) receives the score of each written keypoint (0 in Off mode).
 */
int64_t fdf_oracle_detect(const uint8_t *img, uint32_t w, uint32_t h, uint32_t pitch, uint8_t t,
                          uint8_t n, uint8_t nms, fdf_oracle_point *out, size_t cap,
                          uint16_t *scores_out);

/* SipHash-1-3 (key 0,0) of a point list exactly as tests/compare.rs:5-12 hashes a &[Point]:
 * len as u64 LE, then x:u32 LE, y:u32 LE per point, then Rust's 0xff-less finish. */
uint64_t fdf_oracle_hash_points(const fdf_oracle_point *pts, size_t n);
uint64_t fdf_oracle_siphash13(const uint8_t *data, size_t len);

/*
 * Counter-based synthetic frame generator shared (bit for bit) with the CUDA library's
 * fdf_synth_frames_device: every pixel is a pure function of (seed, frame, x, y, kind, amp).
 *   kind 0: "scene"  - layered random rectangles on a blocky background + uniform noise +-amp
 *   kind 1: "noise"  - uniform random bytes (stress: ~28 % keypoints at t=16 n=9)
 */
void fdf_oracle_synth_frame(uint8_t *out, uint32_t w, uint32_t h, uint32_t pitch, uint64_t seed,
                            uint32_t frame, uint32_t kind, uint32_t amp);

#ifdef __cplusplus
}
#endif
#endif
