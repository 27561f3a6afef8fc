/*
 * fdf.h -- C ABI of libfdf_cuda.so, the B200 (sm_100a) FAST-n corner detector.
 *
 * This is the drop-in boundary for the detection path of iwanders/feature_detector_fast.  Each
 * entry point names the reference interface it replaces (paths into the reference checkout).
 * Plain pointers and sizes only; nothing here unwinds or throws: every call returns an fdf_status.
 *
 * Semantics reproduced bit for bit (see DESIGN.md / SURVEY.md section 3.1): centres
 * x in [3, w-3), y in [3, h-3); brighter <=> p > c+t, darker <=> p < c-t (strict); keypoint <=>
 * a cyclic run of >= count ring pixels all brighter or all darker; NMS keeps a keypoint iff its
 * score is strictly greater than its 8 neighbours' and never emits rows 3 and h-4; output is
 * row-major (y, then x).
 */
#ifndef FDF_H
#define FDF_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* lib.rs:15-20  `pub struct Point { pub x: u32, pub y: u32 }`  (#[repr(C)] mirror) */
typedef struct fdf_point {
    uint32_t x;
    uint32_t y;
} fdf_point;

/* lib.rs:25-36  `pub enum NonMaximalSuppression { Off, MaxThreshold, SumAbsolute }`
 * numbered like fast_simd.rs:74-76 (NONMAX_DISABLED / NONMAX_MAX_THRESHOLD / NONMAX_SUM_ABSOLUTE) */
enum {
    FDF_NMS_OFF = 0,
    FDF_NMS_MAX_THRESHOLD = 1,
    FDF_NMS_SUM_ABSOLUTE = 2
};

typedef enum fdf_status {
    FDF_OK = 0,
    FDF_ERR_INVALID_COUNT = 1,   /* count outside 9..=16: the reference panics (fast_simd.rs:302-305, :797-801) */
    FDF_ERR_INVALID_NMS = 2,     /* nms not one of FDF_NMS_* */
    FDF_ERR_INVALID_ARGUMENT = 3,/* null pointer, pitch < width, misaligned device buffer, ... */
    FDF_ERR_CAPACITY = 4,        /* output buffer too small; *n_out / offsets[n_frames] hold the needed size */
    FDF_ERR_CUDA = 5,            /* a CUDA call failed; see fdf_last_error() */
    FDF_ERR_NO_DEVICE = 6,       /* no usable sm_100 device */
    FDF_ERR_INTERNAL = 7,        /* device-side consistency check failed (a pipeline wait timed out, staging overflow) */
    FDF_ERR_BUSY = 8             /* fdf_pipe_submit: `depth` images are in flight already, collect one first */
} fdf_status;

/* One context per host thread and device: owns a stream, the scan workspace and the staging
 * buffers of the host-memory entry points.  Not thread-safe; create one per thread (the reference
 * function is pure and re-entrant: lib.rs:62-64). */
typedef struct fdf_ctx fdf_ctx;

fdf_status fdf_create(int device, fdf_ctx **out_ctx);
void fdf_destroy(fdf_ctx *ctx);

/*
 * Replaces `feature_detector_fast::detect(img: &GrayImage, config: &Config) -> Vec<Point>`
 * (lib.rs:62-64), `Config::detect` (lib.rs:54-59) and `fast_simd::detector` (fast_simd.rs:847-859).
 *
 * img: HOST memory, h rows of w bytes, `pitch` bytes between rows (GrayImage: pitch == w,
 * fast_simd.rs:307-308, 330).  threshold / count / nms are Config's three fields (lib.rs:38-52).
 * Writes up to `cap` points to the HOST array `out` in the reference's order and the number found
 * to *n_out.  If more than cap were found, returns FDF_ERR_CAPACITY with *n_out = number found
 * (the contents of `out` are then unspecified: call again with room for *n_out points).
 * w < 7 or h < 7 yields 0 points.
 */
fdf_status fdf_detect(fdf_ctx *ctx, const uint8_t *img, uint32_t w, uint32_t h, uint32_t pitch,
                      uint8_t threshold, uint8_t count, uint8_t nms, fdf_point *out, size_t cap,
                      size_t *n_out);

/*
 * The same detector over a batch of n_frames equally sized frames in HOST memory (frame f starts
 * at frames + f * frame_stride).  Output is CSR: the points of frame f are
 * out[offsets[f] .. offsets[f+1]) ; offsets has n_frames + 1 entries, offsets[0] == 0.
 * The timed region of an end-to-end measurement is exactly one call of this function
 * (host->device copy, kernels, device->host copy of offsets and points).
 */
fdf_status fdf_detect_batch(fdf_ctx *ctx, const uint8_t *frames, uint32_t n_frames, uint32_t w,
                            uint32_t h, uint32_t pitch, uint64_t frame_stride, uint8_t threshold,
                            uint8_t count, uint8_t nms, fdf_point *out, size_t cap,
                            uint64_t *offsets);

/*
 * Device-resident, asynchronous form (what a GPU producer upstream of the detector calls):
 * d_frames, d_out and d_offsets are DEVICE pointers; work is enqueued on `stream` (a
 * cudaStream_t; NULL = the CUDA default stream) and the call returns without synchronising.
 * d_frames must be 16-byte aligned with pitch and frame_stride multiples of 16 (TMA tensor-map
 * requirements); otherwise FDF_ERR_INVALID_ARGUMENT.  d_offsets receives n_frames + 1 entries;
 * if offsets[n_frames] > cap the output is incomplete (its contents are unspecified).
 * Uses the context's workspace and staging buffer (cap points): calls on one context must be
 * stream-ordered.
 */
fdf_status fdf_detect_device(fdf_ctx *ctx, const uint8_t *d_frames, uint32_t n_frames, uint32_t w,
                             uint32_t h, uint32_t pitch, uint64_t frame_stride, uint8_t threshold,
                             uint8_t count, uint8_t nms, fdf_point *d_out, size_t cap,
                             uint64_t *d_offsets, void *stream);

/*
 * Multi-GPU (north star / SURVEY 8e): frames are independent units, so a batch of total_frames frames is sharded by
 * contiguous blocks over n_ranks processes, one per GPU -- rank r owns frames [r*total/n, (r+1)*total/n) -- and the
 * path's only exchange step is one all-gather of the ranks' local CSR offsets (NCCL over NVLink; the caller's
 * communication library does it, this library has no networking code).  The reference has no counterpart: its
 * `detect` (lib.rs:62-64) handles one image on one thread.  Per rank and batch:
 *
 *   fdf_detect_device     the rank's frames -> d_points (LOCAL order and offsets) and the n_frames + 1 local offsets,
 *                         written straight into the rank's block of the all-gather buffer.
 *   <all-gather>          every rank's block of `block` >= (largest shard + 1) uint64 entries into
 *                         d_all_offsets[n_ranks * block].
 *   fdf_shard_push        one kernel: copies the rank's points to their place in the batch result,
 *                         d_result[(keypoints of all lower ranks) + i] -- d_result is typically rank 0's buffer mapped
 *                         with fdf_shared_open, so the batch result is assembled over NVLink with coalesced 16-byte
 *                         stores and no host round trip (d_result == NULL: offsets only) -- and writes the batch's global
 *                         CSR offsets (total_frames + 1 entries) to this rank's d_global_offsets.
 * The all-gather and the push do not feed the next batch's detection, so a caller can put them on a second stream
 * and overlap them with it (feature_detector_fast_b200.sharding.ShardedDetector does).  A reader of d_result on another
 * rank needs a later collective or barrier on that stream before it looks at the points.
 */
fdf_status fdf_shard_push(fdf_ctx *ctx, const uint64_t *d_all_offsets, uint32_t block, uint32_t n_ranks, uint32_t rank,
                          uint32_t total_frames, const fdf_point *d_points, fdf_point *d_result, size_t cap_total,
                          uint64_t *d_global_offsets, void *stream);

/* Device memory that other processes on the same node can map (CUDA IPC): fdf_shared_alloc returns the pointer and a
 * 64-byte handle to send to the peers, fdf_shared_open maps a peer's allocation into this process (peer access over
 * NVLink), fdf_shared_close unmaps / frees.  Everything still open is released by fdf_destroy. */
fdf_status fdf_shared_alloc(fdf_ctx *ctx, size_t bytes, void **d_ptr, uint8_t handle[64]);
fdf_status fdf_shared_open(fdf_ctx *ctx, const uint8_t handle[64], void **d_ptr);
fdf_status fdf_shared_close(fdf_ctx *ctx, void *d_ptr);

/*
 * The step in front of the path in the reference's CLI: `image::open(path).to_rgb8()` followed by
 * `DynamicImage::ImageRgb8(..).to_luma8()` (main.rs:53-58).  Interleaved RGB8 in DEVICE memory -> luma8 in
 * DEVICE memory, enqueued on `stream`: luma = (2126 r + 7152 g + 722 b) / 10000, integer, truncating -- the
 * weights of image 0.24.6 (Cargo.lock:388-390; the crate is not vendored in the reference checkout, so this
 * conversion is pinned only by "identity for r == g == b", which the shipped grey PNG confirms).
 * Give luma_pitch / luma_frame_stride as multiples of 16 to pass the result straight to fdf_detect_device.
 */
fdf_status fdf_rgb8_to_luma8_device(fdf_ctx *ctx, const uint8_t *d_rgb, uint32_t n_frames, uint32_t w, uint32_t h,
                                    uint32_t rgb_pitch, uint64_t rgb_frame_stride, uint8_t *d_luma,
                                    uint32_t luma_pitch, uint64_t luma_frame_stride, void *stream);

/*
 * The crate's OWN grey conversion, util.rs:5-41 (`Rgb8ToLuma16View` + `to_grey`; main.rs:57 has it commented out):
 * the view's pixel is r + g + b as u16 (util.rs:38-40) and to_grey stores that / 3 as u8 (util.rs:21-23).  Same
 * layout contract as fdf_rgb8_to_luma8_device.  Unlike the image crate's weights this one is pinned by the
 * reference's source.
 */
fdf_status fdf_rgb8_to_grey_sum3_device(fdf_ctx *ctx, const uint8_t *d_rgb, uint32_t n_frames, uint32_t w, uint32_t h,
                                        uint32_t rgb_pitch, uint64_t rgb_frame_stride, uint8_t *d_grey,
                                        uint32_t grey_pitch, uint64_t grey_frame_stride, void *stream);

/*
 * Streaming form of fdf_detect ("batch / streaming host API with pinned-buffer pool and overlap of
 * H2D || kernel || D2H"; the reference's `detect`, lib.rs:62-64, is synchronous and has no counterpart).  A pipe keeps
 * up to `depth` images in flight on its context: fdf_pipe_submit enqueues host->device copy, the three kernels and
 * the device->host copy of one image on three streams and returns without waiting; fdf_pipe_collect waits for the
 * OLDEST submitted image and hands its keypoints out (first in, first out), exactly what fdf_detect would have
 * returned for it.  Each of the `depth` slots owns pinned host buffers and device buffers for an image of up to
 * max_w x max_h and `cap` keypoints, allocated once in fdf_pipe_create.
 *   - a pageable image is copied into the slot's pinned buffer before submit returns: the caller may reuse it at once;
 *     an image in pinned memory (cudaHostAlloc / cudaHostRegister) is read by the DMA engine directly and must stay
 *     untouched until the matching collect;
 *   - submit with `depth` images in flight returns FDF_ERR_BUSY (nothing is enqueued);
 *   - collect with nothing in flight returns FDF_ERR_INVALID_ARGUMENT; a failed collect (FDF_ERR_CAPACITY: the
 *     caller's `cap`, or the pipe's, is smaller than *n_out; FDF_ERR_INTERNAL) still retires the image.
 * A pipe uses its context's streams and workspace: like every other call on a context it belongs to one host thread,
 * and it must be destroyed before its context.
 */
typedef struct fdf_pipe fdf_pipe;
fdf_status fdf_pipe_create(fdf_ctx *ctx, uint32_t depth, uint32_t max_w, uint32_t max_h, size_t cap, fdf_pipe **out_pipe);
void fdf_pipe_destroy(fdf_pipe *pipe);
fdf_status fdf_pipe_submit(fdf_pipe *pipe, const uint8_t *img, uint32_t w, uint32_t h, uint32_t pitch, uint8_t threshold,
                           uint8_t count, uint8_t nms);
fdf_status fdf_pipe_collect(fdf_pipe *pipe, fdf_point *out, size_t cap, size_t *n_out);
uint32_t fdf_pipe_in_flight(const fdf_pipe *pipe);

/*
 * main.rs:53-67 in one call: an interleaved RGB8 image in HOST memory (h rows of 3 w bytes, rgb_pitch bytes
 * between rows) is copied to the device, converted to luma there and run through the detector.  Same output
 * contract as fdf_detect.
 */
fdf_status fdf_detect_rgb8(fdf_ctx *ctx, const uint8_t *rgb, uint32_t w, uint32_t h, uint32_t rgb_pitch,
                           uint8_t threshold, uint8_t count, uint8_t nms, fdf_point *out, size_t cap,
                           size_t *n_out);

/* Fills device memory with the synthetic frames used by the tests and the benchmark (bit-identical
 * to oracle/fdf_oracle.c: fdf_oracle_synth_frame).  Frame f of the output is generator frame
 * first_frame + f.  kind 0 = scene, 1 = uniform noise; amp = noise amplitude of kind 0. */
fdf_status fdf_synth_frames_device(fdf_ctx *ctx, uint8_t *d_frames, uint32_t n_frames, uint32_t w,
                                   uint32_t h, uint32_t pitch, uint64_t frame_stride, uint64_t seed,
                                   uint32_t first_frame, uint32_t kind, uint32_t amp, void *stream);

/* Number of kernels this context has launched so far (bench.py's gpu_launches): three per detection call
 * (detection, offset scan, gather) and one per synthetic-frame call. */
uint64_t fdf_kernel_launches(const fdf_ctx *ctx);

/* Tuning knobs of a context (tests and experiments; results never depend on them).  strip_rows: scored rows per
 * strip, 0 = chosen from the batch size (default), or 32 / 48 / 64.  sub_batch_mb: how many MB of frames
 * fdf_detect_batch copies and processes at a time, 0 = default (128).  The environment variables FDF_FORCE_SR and
 * FDF_SUB_BATCH_MB give the initial values when the context is created; they are not read afterwards. */
fdf_status fdf_set_tuning(fdf_ctx *ctx, int strip_rows, uint32_t sub_batch_mb);

/* Work items of the detection kernel (tests and experiments; results never depend on it).  The kernel's CTAs draw
 * (frame, strip) tickets.  A small input -- one image: 36 strips for 592 CTA slots -- is cut finer: every strip becomes
 * 2, 4 or 8 equal ranges of chunks, each a ticket of its own.  parts = 0: automatic (default: split while there are
 * fewer items than CTA slots); 1, 2, 4, 8: at most that many.  Only the 32-row-strip kernels (small inputs) split, only
 * evenly and with at least two chunks per item. */
fdf_status fdf_set_item_parts(fdf_ctx *ctx, uint32_t parts);

/* Leave SMs free beside the detection kernel.  That kernel is persistent and fills every SM (three CTAs each, all of
 * the registers and shared memory), so any other kernel that becomes ready while it runs -- the all-gather of a
 * communication library, fdf_shard_push -- waits for the whole launch to drain: measured on 8 GPUs, every exchange then
 * takes one detection period and the ranks run in lock step behind it.  With sm_stride = n > 0 the detection CTAs that
 * land on every n-th SM (%smid % n == n - 1) return at once; the strips are handed out by ticket, so the other CTAs
 * do all the work (throughput -1/n) and the idle SMs take the exchange kernels immediately.  0 = use every SM
 * (default).  Results never depend on it. */
fdf_status fdf_set_idle_sms(fdf_ctx *ctx, uint32_t sm_stride);

/* Per-kernel device timing for benchmarks.  fdf_set_timing(ctx, n) makes every following
 * fdf_detect_device-family call record CUDA events around its three launches into slot
 * (call index mod n); n = 0 switches it off.  fdf_get_timing synchronises on the slot's last event and
 * returns the milliseconds of {detection kernel, offset-scan kernel, gather kernel}. */
fdf_status fdf_set_timing(fdf_ctx *ctx, uint32_t slots);
fdf_status fdf_get_timing(fdf_ctx *ctx, uint32_t slot, float ms[3]);

/* Device-side flags of the last fdf_detect_device-family call on this context, read back with a
 * synchronising copy: 0 = clean, bit 0 = look-back wait timed out, bit 1 = a pipeline wait of the detection kernel
 * timed out, bit 2 = the internal staging buffer overflowed (keypoints were dropped): the result is invalid. */
fdf_status fdf_check_device_flags(fdf_ctx *ctx, uint32_t *flags);

const char *fdf_last_error(const fdf_ctx *ctx);
const char *fdf_status_string(fdf_status status);
const char *fdf_version(void);

#ifdef __cplusplus
}
#endif
#endif /* FDF_H */
