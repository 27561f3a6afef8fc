// fdf_kernels.cuh -- launch interface between the C ABI (fdf_capi.cu) and the kernels (fdf_kernels.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace fdf {

// Geometry shared by host and device.  A frame is cut into STRIPS of full-width rows; a CTA
// takes one strip at a time (atomic ticket) and walks it left to right in CHUNKS.  For every chunk a 256-byte-wide tile
// (chunk + halo) is staged into shared memory by one TMA 3-D tiled load.
constexpr int kComputeWarps = 8;
constexpr int kComputeThreads = kComputeWarps * 32;
constexpr int kThreads = kComputeThreads;              // threads per CTA
constexpr int kTileW = 256;     // tile width in bytes = TMA box inner extent (the maximum)
constexpr int kChunkW = 240;    // output columns per chunk (the last chunk of a row may take one more)
constexpr int kTileLead = 16;   // chunk c's tile starts at image column c*kChunkW - kTileLead: TMA needs the
                                // innermost start coordinate 16-byte aligned (an unaligned one faults)
constexpr int kLeftHalo = 12;   // the chunk's first output column is tile column 12: left halo 12, right halo 4
                                // (4 = 3 ring pixels + 1 NMS neighbour)

constexpr int kTagPeriod = 15;  // score plane entries carry a 4-bit chunk tag (1..15) above the 12-bit score; the
                                // plane is cleared at every strip start and every 15 chunks, so entries of
                                // earlier chunks simply read as "no keypoint" and no per-chunk clear is needed
constexpr int kQueueCap = 2048; // candidate queue entries per chunk (typical fill: ~320 at 64 rows); more -> fallback below
constexpr int kKlistCap = 1024; // confirmed keypoints per chunk the list NMS handles (typical: ~170); more -> dense NMS
constexpr int kGroupRows = 8;   // fallback for dense content: filter kGroupRows x 256 <= kQueueCap centres at a time
constexpr int kWarpQueueCap = 128;  // 16-pixel groups one warp can pass from filter stage 1 to stage 2 per chunk
                                    // (= 32 lanes x 4 rows, the most stage 1 looks at)

__host__ __device__ constexpr int chunks_per_row(int w) { return (w + kChunkW - 1) / kChunkW; }

// scored rows per strip (SR) -> staged rows, emitted rows
__host__ __device__ constexpr int tile_rows(int sr) { return sr + 6; }                             // +-3 ring rows
__host__ __device__ constexpr int out_rows(int mode, int sr) { return mode == 0 ? sr : sr - 2; }  // NMS needs a 1-row score halo
__host__ __device__ constexpr int first_out_row(int mode) { return mode == 0 ? 3 : 4; }          // fast_simd.rs:342 / :589-596

struct DetectParams {
    uint32_t w, h, n_frames;
    uint32_t strips_per_frame;
    uint32_t chunks_per_strip;
    uint32_t words_per_row;  // ceil(w / 32): bit-plane words per row
    uint32_t threshold, count;
    unsigned long long cap;  // capacity of out (and of staging), in points
    uint2 *out;              // fdf_point[cap], packed over the whole batch, row-major per frame
    uint2 *staging;          // fdf_point[cap]: each strip's ordered run at a bump-allocated position
    unsigned long long *offsets;      // n_frames + 1
    unsigned long long *cursor;       // staging bump allocator (zeroed per launch)
    uint32_t *item_count;             // [items] keypoints of each (frame, strip)
    unsigned long long *item_src;     // [items] where the strip's run sits in staging
    unsigned long long *item_dst;     // [items] where it goes in out (exclusive scan of item_count)
    unsigned long long *scan_status;  // look-back words of the scan kernel's tiles (zeroed per launch)
    uint32_t *ticket;                 // strip tickets of the detection kernel (zeroed per launch)
    uint32_t *scan_ticket;            // tile tickets of the scan kernel (zeroed per launch)
    uint32_t *flags;                  // zeroed per launch; bit 0 look-back timeout, bit 1 TMA wait timeout
};

constexpr int kScanThreads = 256;
constexpr int kScanItemsPerThread = 8;
constexpr int kScanTile = kScanThreads * kScanItemsPerThread;  // strips per scan tile

size_t detect_smem_bytes(int mode, int sr, uint32_t words_per_row);

// Enqueues the detection kernel for (mode, sr) on `stream`.  tmap describes the frames as a 3-D
// u8 tensor (x, y, frame) with box (kTileW, tile_rows(sr), 1).
cudaError_t launch_detect(int mode, int sr, const CUtensorMap &tmap, const DetectParams &p, cudaStream_t stream);

// Ordered compaction, off the detection kernel's critical path: an exclusive scan of the per-strip counts
// (single pass, decoupled look-back between scan tiles) and a gather of every strip's run to its final,
// row-major position.
cudaError_t launch_scan(const DetectParams &p, cudaStream_t stream);
cudaError_t launch_gather(const DetectParams &p, cudaStream_t stream);

cudaError_t launch_synth(uint8_t *d_frames, uint32_t n_frames, uint32_t w, uint32_t h, uint32_t pitch,
                         unsigned long long frame_stride, unsigned long long seed, uint32_t first_frame,
                         uint32_t kind, uint32_t amp, cudaStream_t stream);

}  // namespace fdf
