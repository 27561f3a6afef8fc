// fdf_kernels.cuh -- launch interface between the C ABI (fdf_capi.cu) and the kernels (fdf_kernels.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace fdf {

// Geometry shared by host and device.  A frame is cut into STRIPS of full-width rows; a CTA
// takes one strip at a time (atomic ticket) and walks it left to right in CHUNKS.  For every chunk a 256-byte-wide tile
// (chunk + halo) is staged into shared memory by one TMA 3-D tiled load.
constexpr int kFilterWarps = 4;  // warps 0 .. 3 run phase A (dense filter), the others everything per candidate
constexpr int kTestWarps = 6;
constexpr int kThreads = (kFilterWarps + kTestWarps) * 32;  // threads per CTA of the detection kernel
constexpr int kFilterThreads = kFilterWarps * 32;
constexpr int kTestThreads = kTestWarps * 32;
constexpr int kGatherThreads = 256; // threads per CTA of the gather kernel
constexpr int kGatherMaxChunks = kGatherThreads - 1;  // chunks per strip the gather kernel holds records for
constexpr int kTileW = 256;     // tile width in bytes = TMA box inner extent (the maximum)
constexpr int kChunkW = 240;    // output columns per chunk (the last chunk of a row may take one more)
constexpr int kTileLead = 16;   // chunk c's tile starts at image column c*kChunkW - kTileLead: TMA needs the
                                // innermost start coordinate 16-byte aligned (an unaligned one faults)
constexpr int kLeftHalo = 12;   // the chunk's first output column is tile column 12: left halo 12, right halo 4
                                // (4 = 3 ring pixels + 1 NMS neighbour)
constexpr int kPlaneW = 248;    // score plane row pitch in cells: plane column = tile column - kPlaneLead
constexpr int kPlaneLead = 8;   // (scored tile columns are 11 .. 252 = plane columns 3 .. 244; the NMS reads 3 .. 245)

constexpr int kTagPeriod = 15;  // score plane entries carry a 4-bit chunk tag (1..15) above the 12-bit score; the
                                // plane is cleared every 15 chunks a CTA processes, so entries of
                                // earlier chunks simply read as "no keypoint" and no per-chunk clear is needed
constexpr int kQueueCap = 1024; // candidate queue entries per chunk (typical fill: ~320 at 64 rows); more -> dense path.
                                // The keypoint lists have the same capacity, so they cannot overflow when the queue did not.
constexpr int kStageBlock = 4096;   // staging entries a CTA reserves at a time from the global cursor
constexpr int kWarpQueueCap = 256;  // 16-pixel groups one warp can pass from filter stage 1 to stage 2 per chunk
                                    // (= 32 lanes x 8 rows, the most stage 1 looks at)

__host__ __device__ constexpr int chunks_per_row(int w) { return (w + kChunkW - 1) / kChunkW; }

// scored rows per strip (SR) -> staged rows, emitted rows
__host__ __device__ constexpr int tile_rows(int sr) { return sr + 6; }                             // +-3 ring rows
__host__ __device__ constexpr int out_rows(int mode, int sr) { return mode == 0 ? sr : sr - 2; }  // NMS needs a 1-row score halo
__host__ __device__ constexpr int first_out_row(int mode) { return mode == 0 ? 3 : 4; }          // fast_simd.rs:342 / :589-596

struct DetectParams {
    uint32_t w, h, n_frames;
    uint32_t strips_per_frame;
    uint32_t chunks_per_strip;
    uint32_t parts;          // work items per strip of the detection kernel (1, 2, 4 or 8 equal chunk ranges; divides chunks_per_strip)
    uint32_t words_per_row;  // ceil(w / 32): bit-plane words per row (gather kernel)
    uint32_t threshold, count;
    uint32_t mode, sr;       // (the gather kernel is not templated)
    uint32_t idle_sm_stride; // > 0: detection CTAs on SMs with %smid % stride == stride - 1 return at once (fdf_set_idle_sms)
    unsigned long long cap;  // capacity of out, in points
    unsigned long long staging_cap;   // capacity of staging: 2 cap + one block per CTA (blocks are not used to the end)
    uint2 *out;              // fdf_point[cap], packed over the whole batch, row-major per frame
    uint32_t *staging;       // [staging_cap] keypoints as (row in strip << 16 | x), one unordered run per chunk
    unsigned long long *offsets;      // n_frames + 1
    unsigned long long *cursor;       // staging bump allocator (zeroed per launch)
    uint32_t *item_count;             // [items] keypoints of each (frame, strip)
    unsigned long long *item_dst;     // [items] where the strip's points go in out (exclusive scan of item_count)
    unsigned long long *run_base;     // [items * chunks] where the chunk's run sits in staging
    uint32_t *run_count;              // [items * chunks] its length (every record is written; 0 = no run)
    unsigned long long *scan_status;  // look-back words of the scan kernel's tiles (zeroed per launch)
    uint32_t *ticket;                 // strip tickets of the detection kernel (zeroed per launch)
    uint32_t *scan_ticket;            // tile tickets of the scan kernel (zeroed per launch)
    uint32_t *flags_copy;             // optional: the gather kernel copies *flags here (single-image path: the flags
                                      // then come home with the offsets and points in one copy)
    uint32_t *flags;                  // zeroed per launch; bit 0 look-back timeout, bit 1 TMA wait timeout,
                                      // bit 2 staging buffer overflow (entries dropped)
};

// Sharded batches (multi-GPU, one process per GPU): frames [lo, hi) of a batch of `total` frames owned by rank r of n
// (contiguous blocks; sharding.frame_shard)
__host__ __device__ inline uint32_t shard_lo(uint32_t total, uint32_t r, uint32_t n) {
    return (uint32_t)(((unsigned long long)r * total) / n);
}

constexpr int kScanThreads = 256;
constexpr int kScanItemsPerThread = 8;
constexpr int kScanTile = kScanThreads * kScanItemsPerThread;  // strips per scan tile

// What fdf_create looks up once about the device and the kernels (see init_device_info).
struct DeviceInfo {
    int sms = 0;
    int detect_per_sm[3][3] = {};       // resident CTAs per SM of fdf_detect_kernel<mode, 32 | 48 | 64>
    size_t gather_smem = ~(size_t)0;    // dynamic shared memory of the last gather launch ...
    int gather_per_sm = 0;              // ... and the gather kernel's occupancy at that size
    int ctas_limit = 0;                 // FDF_CTAS_PER_SM (experiments): cap on detect CTAs per SM, 0 = none
};
cudaError_t init_device_info(DeviceInfo &info);  // (on the current device)

size_t detect_smem_bytes(int mode, int sr);
size_t gather_smem_bytes(int mode, int sr, uint32_t words_per_row);
bool gather_scans_itself(const DetectParams &p);  // few strips: no scan launch, the gather kernel scans the counts

// Enqueues the detection kernel for (mode, sr) on `stream`.  tmap describes the frames as a 3-D
// u8 tensor (x, y, frame) with box (kTileW, tile_rows(sr), 1).
cudaError_t launch_detect(int mode, int sr, const CUtensorMap &tmap, const DetectParams &p, cudaStream_t stream,
                          const DeviceInfo &info);

// Ordered compaction, off the detection kernel's critical path: an exclusive scan of the per-strip counts
// (single pass, decoupled look-back between scan tiles), then one CTA per strip turns the strip's unordered
// runs into row-major points at their final position (through a bit plane of the strip in shared memory).
cudaError_t launch_scan(const DetectParams &p, cudaStream_t stream);
cudaError_t launch_gather(const DetectParams &p, cudaStream_t stream, DeviceInfo &info);

// Sharded batches, after the all-gather of the ranks' local CSR offsets (all_offsets: rank r's block starts at
// r * block and has frames(r) + 1 entries): copies this rank's points[0 .. its total) to result[(sum of the lower ranks'
// totals) + i] -- `result` may be another GPU's memory mapped over NVLink; the copy is coalesced, 16 bytes per thread --
// and writes the batch's global CSR offsets (total_frames + 1 entries).
cudaError_t launch_shard_push(const unsigned long long *all_offsets, uint32_t block, uint32_t n_ranks, uint32_t rank,
                              uint32_t total_frames, const uint2 *points, uint2 *result, unsigned long long cap_total,
                              unsigned long long *global_offsets, int sms, cudaStream_t stream);

// RGB8 (interleaved, 3 bytes per pixel) -> luma8: kind 0 = the `image` crate's integer weights (main.rs:53-58),
// kind 1 = (r + g + b) / 3 as util.rs:5-41 (`Rgb8ToLuma16View::to_grey`) does.
cudaError_t launch_luma(const uint8_t *d_rgb, uint32_t n_frames, uint32_t w, uint32_t h, uint32_t rgb_pitch,
                        unsigned long long rgb_stride, uint8_t *d_luma, uint32_t luma_pitch,
                        unsigned long long luma_stride, int kind, cudaStream_t stream);

cudaError_t launch_synth(uint8_t *d_frames, uint32_t n_frames, uint32_t w, uint32_t h, uint32_t pitch,
                         unsigned long long frame_stride, unsigned long long seed, uint32_t first_frame,
                         uint32_t kind, uint32_t amp, cudaStream_t stream);

}  // namespace fdf
