// fdf_capi.cu -- the C ABI of libfdf_cuda.so (include/fdf.h): context, argument checking, TMA
// tensor-map encoding, staging for the host-memory entry points.  No torch types, no CPU fallback:
// every entry point either runs the sm_100a kernels or returns an error status.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/fdf.h"
#include "fdf_kernels.cuh"

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <typename T>
struct DeviceBuffer {
    T *ptr = nullptr;
    size_t count = 0;
    cudaError_t reserve(size_t n) {
        if (n <= count) return cudaSuccess;
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        count = 0;
        cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&ptr), n * sizeof(T));
        if (e == cudaSuccess) count = n;
        return e;
    }
    void release() {
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        count = 0;
    }
};

}  // namespace

#ifdef FDF_TRACE
namespace fdf {
cudaError_t read_trace(long long *out, size_t bytes);
}
#endif
#ifdef FDF_CHECKS
namespace fdf {
cudaError_t read_check_failure(int *line);
}
#endif

struct fdf_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;       // kernels of the host-memory entry points
    cudaStream_t copy_stream = nullptr;  // their host->device copies (overlap the kernels of earlier sub-batches)
    cudaStream_t back_stream = nullptr;  // their device->host copies of points
    std::vector<cudaEvent_t> pipe_events;  // per sub-batch: frames landed, kernels done
    EncodeTiledFn encode = nullptr;
    DeviceBuffer<uint8_t> workspace;            // tickets | flags | cursor | scan status | per-strip count/dst | run records
    DeviceBuffer<uint32_t> staging;             // per-chunk unordered runs of (row << 16 | x) before the gather
    DeviceBuffer<uint8_t> staged_frames;        // host-path input staging (pitched to 16 bytes)
    DeviceBuffer<uint8_t> staged_rgb;           // fdf_detect_rgb8: the RGB image before the luma conversion
    DeviceBuffer<fdf_point> staged_points;      // host-path output staging
    DeviceBuffer<unsigned long long> staged_offsets;
    unsigned long long *pinned_offsets = nullptr;
    size_t pinned_offsets_count = 0;
    // single-image path (fdf_detect): one device block [offsets u64 x 2 | flags | pad to 64 B | points], a pinned copy of
    // its head for the one device->host copy of a call, and a pinned staging buffer for pageable input images
    DeviceBuffer<uint8_t> single_block;
    uint8_t *pinned_head = nullptr;
    size_t pinned_head_bytes = 0;
    uint8_t *pinned_in = nullptr;
    size_t pinned_in_bytes = 0;
    uint32_t *flags_copy_next = nullptr;  // fdf_detect -> fdf_detect_device: where the next launch copies its flags
    size_t single_hint = 0;               // keypoints of the previous fdf_detect call (sizes the first copy back)
    // the tensor map of the previous launch (cuTensorMapEncodeTiled costs a microsecond or two per call)
    struct TmapKey {
        const void *base = nullptr;
        uint32_t w = 0, h = 0, n = 0, pitch = 0;
        uint64_t stride = 0;
        int sr = 0;
        bool operator==(const TmapKey &o) const {
            return base == o.base && w == o.w && h == o.h && n == o.n && pitch == o.pitch && stride == o.stride && sr == o.sr;
        }
    } tmap_key;
    CUtensorMap tmap_cached;
    uint64_t launches = 0;
    fdf::DeviceInfo info;        // SM count, kernel occupancies, experiment knobs: looked up once in fdf_create
    uint32_t item_parts = 0;     // fdf_set_item_parts: work items per strip, 0 = chosen from the batch size
    uint32_t idle_sm_stride = 0; // fdf_set_idle_sms: every n-th SM is left to other kernels by the detection kernel
    int force_sr = 0;            // FDF_FORCE_SR (experiments / tests): strip height override, read once in fdf_create
    unsigned long long sub_batch_bytes = 128ull << 20;  // fdf_detect_batch sub-batch size (FDF_SUB_BATCH_MB, read once)
    std::vector<void *> shared_owned, shared_opened;  // fdf_shared_alloc / fdf_shared_open
    std::vector<cudaEvent_t> timing_events;  // 4 per slot: before detection, after it, after scan, after gather
    uint64_t timing_calls = 0;
    char error[512] = {0};
};

namespace {

constexpr size_t kWorkspaceHeader = 64;  // ticket at +0, flags at +4, status words from +64

fdf_status fail(fdf_ctx *ctx, fdf_status st, const char *fmt, ...) {
    if (ctx) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(ctx->error, sizeof(ctx->error), fmt, ap);
        va_end(ap);
    }
    return st;
}

#define FDF_CUDA(ctx, call)                                                                          \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess)                                                                       \
            return fail(ctx, FDF_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, \
                        __LINE__);                                                                   \
    } while (0)

fdf_status check_config(fdf_ctx *ctx, uint8_t count, uint8_t nms) {
    // fast_simd.rs:302-305 asserts count >= 9; count > 16 panics at :797-801
    if (count < 9 || count > 16) return fail(ctx, FDF_ERR_INVALID_COUNT, "count must be in 9..=16, got %u", count);
    if (nms > FDF_NMS_SUM_ABSOLUTE) return fail(ctx, FDF_ERR_INVALID_NMS, "unknown nms mode %u", nms);
    return FDF_OK;
}

// strip height: tall strips (less halo) when there is enough work to fill the GPU, short otherwise
int choose_scored_rows(const fdf_ctx *ctx, uint32_t n_frames, uint32_t h, int mode) {
    const long long rows = (long long)h - 2 * fdf::first_out_row(mode);
    if (ctx->force_sr == 32 || ctx->force_sr == 48 || ctx->force_sr == 64) return ctx->force_sr;
    const long long strips64 = (rows + fdf::out_rows(mode, 64) - 1) / fdf::out_rows(mode, 64);
    return (long long)n_frames * strips64 >= 2 * 148 ? 64 : 32;
}

}  // namespace

extern "C" {

const char *fdf_version(void) { return "fdf-b200 0.1 (sm_100a)"; }

const char *fdf_status_string(fdf_status status) {
    switch (status) {
        case FDF_OK: return "ok";
        case FDF_ERR_INVALID_COUNT: return "count must be in 9..=16";
        case FDF_ERR_INVALID_NMS: return "unknown non-maximal-suppression mode";
        case FDF_ERR_INVALID_ARGUMENT: return "invalid argument";
        case FDF_ERR_CAPACITY: return "output capacity too small";
        case FDF_ERR_CUDA: return "CUDA error";
        case FDF_ERR_NO_DEVICE: return "no usable sm_100 device";
        case FDF_ERR_INTERNAL: return "device-side consistency check failed";
        case FDF_ERR_BUSY: return "pipe full: collect an image first";
    }
    return "unknown status";
}

const char *fdf_last_error(const fdf_ctx *ctx) { return ctx ? ctx->error : "null context"; }

uint64_t fdf_kernel_launches(const fdf_ctx *ctx) { return ctx ? ctx->launches : 0; }

fdf_status fdf_create(int device, fdf_ctx **out_ctx) {
    if (!out_ctx) return FDF_ERR_INVALID_ARGUMENT;
    *out_ctx = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return FDF_ERR_NO_DEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return FDF_ERR_NO_DEVICE;
    if (prop.major != 10) return FDF_ERR_NO_DEVICE;  // the only code in this library is sm_100a SASS
    fdf_ctx *ctx = new (std::nothrow) fdf_ctx();
    if (!ctx) return FDF_ERR_INTERNAL;
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->back_stream, cudaStreamNonBlocking) != cudaSuccess) {
        fdf_destroy(ctx);  // (frees whichever streams were created)
        return FDF_ERR_CUDA;
    }
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
        qres != cudaDriverEntryPointSuccess) {
        fdf_destroy(ctx);
        return FDF_ERR_CUDA;
    }
    ctx->encode = reinterpret_cast<EncodeTiledFn>(fn);
    if (fdf::init_device_info(ctx->info) != cudaSuccess) {
        fdf_destroy(ctx);
        return FDF_ERR_CUDA;
    }
    // tuning / test knobs are read here, once: the environment cannot change the behaviour of a live context
    if (const char *force = getenv("FDF_FORCE_SR")) ctx->force_sr = atoi(force);
    if (const char *mb = getenv("FDF_SUB_BATCH_MB")) {
        const long v = atol(mb);
        if (v >= 1 && v <= 65536) ctx->sub_batch_bytes = (unsigned long long)v << 20;
    }
    *out_ctx = ctx;
    return FDF_OK;
}

void fdf_destroy(fdf_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (cudaEvent_t e : ctx->timing_events) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->pipe_events) cudaEventDestroy(e);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->back_stream) cudaStreamDestroy(ctx->back_stream);
    for (void *q : ctx->shared_opened) cudaIpcCloseMemHandle(q);
    for (void *q : ctx->shared_owned) cudaFree(q);
    ctx->workspace.release();
    ctx->staging.release();
    ctx->staged_frames.release();
    ctx->staged_rgb.release();
    ctx->staged_points.release();
    ctx->staged_offsets.release();
    if (ctx->pinned_offsets) cudaFreeHost(ctx->pinned_offsets);
    if (ctx->pinned_head) cudaFreeHost(ctx->pinned_head);
    if (ctx->pinned_in) cudaFreeHost(ctx->pinned_in);
    ctx->single_block.release();
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

namespace {
// Argument checks, workspace / staging sizing and the TMA tensor map of one detection call.  Returns FDF_OK with
// *empty = true when nothing can be a keypoint (the caller then only zeroes its offsets).
fdf_status prepare_detect(fdf_ctx *ctx, const uint8_t *d_frames, uint32_t n_frames, uint32_t w, uint32_t h, uint32_t pitch,
                          uint64_t frame_stride, uint8_t threshold, uint8_t count, uint8_t nms, size_t cap,
                          cudaStream_t stream, fdf::DetectParams &p, CUtensorMap &tmap, bool *empty) {
    *empty = false;
    const int mode = nms;
    const long long rows = (long long)h - 2 * fdf::first_out_row(mode);
    if (n_frames == 0 || w < 7 || h < 7 || rows <= 0) {  // nothing can be a keypoint
        *empty = true;
        return FDF_OK;
    }
    if (!d_frames) return fail(ctx, FDF_ERR_INVALID_ARGUMENT, "null frame pointer");
    if (pitch < w) return fail(ctx, FDF_ERR_INVALID_ARGUMENT, "pitch %u < width %u", pitch, w);
    if (n_frames > 1 && frame_stride < (uint64_t)pitch * h)
        return fail(ctx, FDF_ERR_INVALID_ARGUMENT, "frame_stride smaller than one frame");
    if (n_frames == 1) frame_stride = (((uint64_t)pitch * h) + 15ull) & ~15ull;
    if ((reinterpret_cast<uintptr_t>(d_frames) & 15u) || (pitch & 15u) || (frame_stride & 15ull))
        return fail(ctx, FDF_ERR_INVALID_ARGUMENT,
                    "device frames need a 16-byte aligned base, pitch and frame_stride (TMA tensor map)");

    const int sr = choose_scored_rows(ctx, n_frames, h, mode);
    p = fdf::DetectParams();
    p.w = w;
    p.h = h;
    p.n_frames = n_frames;
    p.strips_per_frame = (uint32_t)((rows + fdf::out_rows(mode, sr) - 1) / fdf::out_rows(mode, sr));
    p.chunks_per_strip = (uint32_t)fdf::chunks_per_row((int)w);
    p.words_per_row = (w + 31) / 32;
    p.threshold = threshold;
    p.count = count;
    p.mode = (uint32_t)mode;
    p.sr = (uint32_t)sr;
    p.idle_sm_stride = ctx->idle_sm_stride;
    p.cap = cap;

    if (fdf::gather_smem_bytes(mode, sr, p.words_per_row) > 200 * 1024 || w > 65535u ||
        p.chunks_per_strip > (uint32_t)fdf::kGatherMaxChunks)
        return fail(ctx, FDF_ERR_INVALID_ARGUMENT, "image too wide (%u) for the strip bit plane", w);
    const unsigned long long items = (unsigned long long)n_frames * p.strips_per_frame;
    if (items > 0x7fffffffull) return fail(ctx, FDF_ERR_INVALID_ARGUMENT, "batch too large");
    // Work items of the detection kernel: whole strips, except for small inputs (the 32-row kernels: fewer than 2 x 148
    // strips of 64 rows in the batch), where every strip is cut into 2, 4 or 8 equal chunk ranges while that still
    // raises the number of busy CTAs -- one 1080p image is 36 strips, i.e. 36 of 592 CTA slots and a chain of 8 chunks
    // per CTA.  Only even splits, at least two chunks (= tile stages) per item.
    {
        const unsigned long long slots = (unsigned long long)ctx->info.sms * 4ull;
        const uint32_t want = ctx->item_parts ? ctx->item_parts : 8u;
        uint32_t parts = 1u;
        while (sr == 32 && 2u * parts <= want && (ctx->item_parts != 0u || items * parts < slots) &&
               p.chunks_per_strip % (2u * parts) == 0u && p.chunks_per_strip / (2u * parts) >= 2u)
            parts *= 2u;
        p.parts = parts;
    }

    // workspace: [header 64 B: ticket, flags, scan ticket, cursor][scan status][zeroed up to here per launch]
    //            [item_dst][run_base][item_count][run_count]
    const size_t scan_tiles = ((size_t)items + fdf::kScanTile - 1) / fdf::kScanTile;
    const size_t zeroed_bytes = kWorkspaceHeader + scan_tiles * sizeof(unsigned long long);
    const size_t runs = (size_t)items * p.chunks_per_strip;  // one run record per chunk
    const size_t dst_off = (zeroed_bytes + 15) & ~(size_t)15;
    const size_t rbase_off = dst_off + (size_t)items * sizeof(unsigned long long);
    const size_t cnt_off = rbase_off + runs * sizeof(unsigned long long);
    const size_t rcnt_off = cnt_off + (size_t)items * sizeof(uint32_t);
    const size_t ws_bytes = rcnt_off + runs * sizeof(uint32_t);
    FDF_CUDA(ctx, ctx->workspace.reserve(ws_bytes));
    // Staging holds one unordered run per chunk, cut from kStageBlock-entry blocks that a CTA takes from a global
    // cursor.  A block's unused tail is lost when the next run may not fit it (dense content reserves exact sizes and
    // may leave most of a block behind), so K keypoints can take up to 2 K entries plus one partly used block per CTA
    // of the grid.  Overflow is not silent: the kernels raise flag bit 2.
    {
        unsigned long long ctas = (unsigned long long)ctx->info.sms * 4ull;  // (at most 4 resident CTAs per SM)
        if (ctas > items * p.parts) ctas = items * p.parts;  // (every CTA that gets a ticket opens a block)
        p.staging_cap = 2ull * cap + (ctas + 1ull) * (unsigned long long)fdf::kStageBlock;
    }
    FDF_CUDA(ctx, ctx->staging.reserve((size_t)p.staging_cap));
    FDF_CUDA(ctx, cudaMemsetAsync(ctx->workspace.ptr, 0, zeroed_bytes, stream));
    p.ticket = reinterpret_cast<uint32_t *>(ctx->workspace.ptr);
    p.flags = reinterpret_cast<uint32_t *>(ctx->workspace.ptr + 4);
    p.scan_ticket = reinterpret_cast<uint32_t *>(ctx->workspace.ptr + 8);
    p.cursor = reinterpret_cast<unsigned long long *>(ctx->workspace.ptr + 16);
    p.scan_status = reinterpret_cast<unsigned long long *>(ctx->workspace.ptr + kWorkspaceHeader);
    p.item_dst = reinterpret_cast<unsigned long long *>(ctx->workspace.ptr + dst_off);
    p.run_base = reinterpret_cast<unsigned long long *>(ctx->workspace.ptr + rbase_off);
    p.item_count = reinterpret_cast<uint32_t *>(ctx->workspace.ptr + cnt_off);
    p.run_count = reinterpret_cast<uint32_t *>(ctx->workspace.ptr + rcnt_off);
    p.staging = ctx->staging.ptr;
    if (p.parts > 1u)  // the items of a strip add their counts up
        FDF_CUDA(ctx, cudaMemsetAsync(p.item_count, 0, (size_t)items * sizeof(uint32_t), stream));

    // frames as a 3-D u8 tensor (x, y, frame); box = one tile; out-of-bounds elements read as 0
    const cuuint64_t dims[3] = {w, h, n_frames};
    const cuuint64_t strides[2] = {pitch, frame_stride};
    const cuuint32_t box[3] = {(cuuint32_t)fdf::kTileW, (cuuint32_t)fdf::tile_rows(sr), 1u};
    const cuuint32_t elem_strides[3] = {1u, 1u, 1u};
    fdf_ctx::TmapKey key;
    key.base = d_frames, key.w = w, key.h = h, key.n = n_frames, key.pitch = pitch, key.stride = frame_stride, key.sr = sr;
    if (key == ctx->tmap_key) {
        tmap = ctx->tmap_cached;
        return FDF_OK;
    }
    CUresult cr = ctx->encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t *>(d_frames), dims, strides,
                              box, elem_strides, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return fail(ctx, FDF_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)cr);
    ctx->tmap_key = key;
    ctx->tmap_cached = tmap;
    return FDF_OK;
}
}  // namespace

fdf_status fdf_detect_device(fdf_ctx *ctx, const uint8_t *d_frames, uint32_t n_frames, uint32_t w, uint32_t h,
                             uint32_t pitch, uint64_t frame_stride, uint8_t threshold, uint8_t count, uint8_t nms,
                             fdf_point *d_out, size_t cap, uint64_t *d_offsets, void *stream_handle) {
    if (!ctx) return FDF_ERR_INVALID_ARGUMENT;
    fdf_status st = check_config(ctx, count, nms);
    if (st != FDF_OK) return st;
    if (!d_offsets || (!d_out && cap > 0)) return fail(ctx, FDF_ERR_INVALID_ARGUMENT, "null output pointer");
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_handle);  // NULL = the default stream
    FDF_CUDA(ctx, cudaSetDevice(ctx->device));
    fdf::DetectParams p;
    CUtensorMap tmap;
    bool empty = false;
    st = prepare_detect(ctx, d_frames, n_frames, w, h, pitch, frame_stride, threshold, count, nms, cap, stream, p, tmap,
                        &empty);
    if (st != FDF_OK) return st;
    uint32_t *flags_copy = ctx->flags_copy_next;
    ctx->flags_copy_next = nullptr;
    if (empty) {
        FDF_CUDA(ctx, cudaMemsetAsync(d_offsets, 0, ((size_t)n_frames + 1) * sizeof(uint64_t), stream));
        if (flags_copy) FDF_CUDA(ctx, cudaMemsetAsync(flags_copy, 0, sizeof(uint32_t), stream));
        return FDF_OK;
    }
    p.out = reinterpret_cast<uint2 *>(d_out);
    p.offsets = reinterpret_cast<unsigned long long *>(d_offsets);
    p.flags_copy = flags_copy;

    cudaEvent_t *ev = nullptr;
    if (!ctx->timing_events.empty()) {
        const size_t slots = ctx->timing_events.size() / 4;
        ev = &ctx->timing_events[4 * (size_t)(ctx->timing_calls++ % slots)];
        FDF_CUDA(ctx, cudaEventRecord(ev[0], stream));
    }
    FDF_CUDA(ctx, fdf::launch_detect((int)p.mode, (int)p.sr, tmap, p, stream, ctx->info));
    if (ev) FDF_CUDA(ctx, cudaEventRecord(ev[1], stream));
    const bool own_scan = fdf::gather_scans_itself(p);  // (one image or a handful: one launch less)
    if (!own_scan) FDF_CUDA(ctx, fdf::launch_scan(p, stream));
    if (ev) FDF_CUDA(ctx, cudaEventRecord(ev[2], stream));
    FDF_CUDA(ctx, fdf::launch_gather(p, stream, ctx->info));
    if (ev) FDF_CUDA(ctx, cudaEventRecord(ev[3], stream));
    ctx->launches += own_scan ? 2 : 3;  // detection, scan, gather
    return FDF_OK;
}

// ---- sharded batches: one process per GPU, one exchange step -------------------------------------------
fdf_status fdf_shard_push(fdf_ctx *ctx, const uint64_t *d_all_offsets, uint32_t block, uint32_t n_ranks, uint32_t rank,
                          uint32_t total_frames, const fdf_point *d_points, fdf_point *d_result, size_t cap_total,
                          uint64_t *d_global_offsets, void *stream_handle) {
    if (!ctx) return FDF_ERR_INVALID_ARGUMENT;
    if (!d_all_offsets || !d_global_offsets || (d_result && !d_points))
        return fail(ctx, FDF_ERR_INVALID_ARGUMENT, "null pointer");
    if (n_ranks == 0 || rank >= n_ranks) return fail(ctx, FDF_ERR_INVALID_ARGUMENT, "bad rank %u of %u", rank, n_ranks);
    uint32_t widest = 0;
    for (uint32_t r = 0; r < n_ranks; r++) {
        const uint32_t fr = fdf::shard_lo(total_frames, r + 1, n_ranks) - fdf::shard_lo(total_frames, r, n_ranks);
        if (fr > widest) widest = fr;
    }
    if (block < widest + 1) return fail(ctx, FDF_ERR_INVALID_ARGUMENT, "block %u < largest shard + 1 (%u)", block, widest + 1);
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_handle);
    FDF_CUDA(ctx, cudaSetDevice(ctx->device));
    FDF_CUDA(ctx, fdf::launch_shard_push(reinterpret_cast<const unsigned long long *>(d_all_offsets), block, n_ranks, rank,
                                         total_frames, reinterpret_cast<const uint2 *>(d_points),
                                         reinterpret_cast<uint2 *>(d_result), cap_total,
                                         reinterpret_cast<unsigned long long *>(d_global_offsets), ctx->info.sms, stream));
    ctx->launches += 1;
    return FDF_OK;
}

// ---- device memory that the other ranks' processes can map (CUDA IPC): the assembled batch result on rank 0 ------
fdf_status fdf_shared_alloc(fdf_ctx *ctx, size_t bytes, void **d_ptr, uint8_t handle[64]) {
    if (!ctx || !d_ptr || !handle) return FDF_ERR_INVALID_ARGUMENT;
    *d_ptr = nullptr;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "fdf.h promises a 64-byte handle");
    FDF_CUDA(ctx, cudaSetDevice(ctx->device));
    void *ptr = nullptr;
    FDF_CUDA(ctx, cudaMalloc(&ptr, bytes ? bytes : 1));
    cudaIpcMemHandle_t hd;
    cudaError_t e = cudaIpcGetMemHandle(&hd, ptr);
    if (e != cudaSuccess) {
        cudaFree(ptr);
        return fail(ctx, FDF_ERR_CUDA, "cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    }
    memcpy(handle, &hd, 64);
    ctx->shared_owned.push_back(ptr);
    *d_ptr = ptr;
    return FDF_OK;
}

fdf_status fdf_shared_open(fdf_ctx *ctx, const uint8_t handle[64], void **d_ptr) {
    if (!ctx || !d_ptr || !handle) return FDF_ERR_INVALID_ARGUMENT;
    *d_ptr = nullptr;
    FDF_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t hd;
    memcpy(&hd, handle, 64);
    void *ptr = nullptr;
    FDF_CUDA(ctx, cudaIpcOpenMemHandle(&ptr, hd, cudaIpcMemLazyEnablePeerAccess));
    ctx->shared_opened.push_back(ptr);
    *d_ptr = ptr;
    return FDF_OK;
}

fdf_status fdf_shared_close(fdf_ctx *ctx, void *d_ptr) {
    if (!ctx || !d_ptr) return FDF_ERR_INVALID_ARGUMENT;
    FDF_CUDA(ctx, cudaSetDevice(ctx->device));
    for (size_t i = 0; i < ctx->shared_owned.size(); i++)
        if (ctx->shared_owned[i] == d_ptr) {
            ctx->shared_owned.erase(ctx->shared_owned.begin() + (long)i);
            FDF_CUDA(ctx, cudaFree(d_ptr));
            return FDF_OK;
        }
    for (size_t i = 0; i < ctx->shared_opened.size(); i++)
        if (ctx->shared_opened[i] == d_ptr) {
            ctx->shared_opened.erase(ctx->shared_opened.begin() + (long)i);
            FDF_CUDA(ctx, cudaIpcCloseMemHandle(d_ptr));
            return FDF_OK;
        }
    return fail(ctx, FDF_ERR_INVALID_ARGUMENT, "pointer was not obtained from fdf_shared_alloc / fdf_shared_open");
}

fdf_status fdf_set_tuning(fdf_ctx *ctx, int strip_rows, uint32_t sub_batch_mb) {
    if (!ctx) return FDF_ERR_INVALID_ARGUMENT;
    if (strip_rows != 0 && strip_rows != 32 && strip_rows != 48 && strip_rows != 64)
        return fail(ctx, FDF_ERR_INVALID_ARGUMENT, "strip_rows must be 0 (automatic), 32, 48 or 64");
    if (sub_batch_mb > 65536u) return fail(ctx, FDF_ERR_INVALID_ARGUMENT, "sub_batch_mb out of range");
    ctx->force_sr = strip_rows;
    ctx->sub_batch_bytes = (unsigned long long)(sub_batch_mb ? sub_batch_mb : 128u) << 20;
    return FDF_OK;
}

fdf_status fdf_set_item_parts(fdf_ctx *ctx, uint32_t parts) {
    if (!ctx) return FDF_ERR_INVALID_ARGUMENT;
    if (parts != 0u && parts != 1u && parts != 2u && parts != 4u && parts != 8u)
        return fail(ctx, FDF_ERR_INVALID_ARGUMENT, "parts must be 0 (automatic), 1, 2, 4 or 8");
    ctx->item_parts = parts;
    return FDF_OK;
}

fdf_status fdf_set_idle_sms(fdf_ctx *ctx, uint32_t sm_stride) {
    if (!ctx) return FDF_ERR_INVALID_ARGUMENT;
    if (sm_stride == 1u) return fail(ctx, FDF_ERR_INVALID_ARGUMENT, "sm_stride 1 would leave no SM to the detection kernel");
    ctx->idle_sm_stride = sm_stride;
    return FDF_OK;
}

fdf_status fdf_set_timing(fdf_ctx *ctx, uint32_t slots) {
    if (!ctx) return FDF_ERR_INVALID_ARGUMENT;
    FDF_CUDA(ctx, cudaSetDevice(ctx->device));
    for (cudaEvent_t e : ctx->timing_events) cudaEventDestroy(e);
    ctx->timing_events.clear();
    ctx->timing_calls = 0;
    for (uint32_t i = 0; i < 4 * slots; i++) {
        cudaEvent_t e;
        FDF_CUDA(ctx, cudaEventCreate(&e));
        ctx->timing_events.push_back(e);
    }
    return FDF_OK;
}

fdf_status fdf_get_timing(fdf_ctx *ctx, uint32_t slot, float ms[3]) {
    if (!ctx || !ms) return FDF_ERR_INVALID_ARGUMENT;
    if ((size_t)slot * 4 + 3 >= ctx->timing_events.size()) return fail(ctx, FDF_ERR_INVALID_ARGUMENT, "no such timing slot");
    cudaEvent_t *ev = &ctx->timing_events[4 * (size_t)slot];
    FDF_CUDA(ctx, cudaEventSynchronize(ev[3]));
    for (int k = 0; k < 3; k++) FDF_CUDA(ctx, cudaEventElapsedTime(&ms[k], ev[k], ev[k + 1]));
    return FDF_OK;
}

fdf_status fdf_check_device_flags(fdf_ctx *ctx, uint32_t *flags) {
    if (!ctx || !flags) return FDF_ERR_INVALID_ARGUMENT;
    *flags = 0;
    if (!ctx->workspace.ptr) return FDF_OK;
    FDF_CUDA(ctx, cudaSetDevice(ctx->device));
    FDF_CUDA(ctx, cudaDeviceSynchronize());
    FDF_CUDA(ctx, cudaMemcpy(flags, ctx->workspace.ptr + 4, sizeof(uint32_t), cudaMemcpyDeviceToHost));
    return FDF_OK;
}

fdf_status fdf_detect_batch(fdf_ctx *ctx, const uint8_t *frames, uint32_t n_frames, uint32_t w, uint32_t h,
                            uint32_t pitch, uint64_t frame_stride, uint8_t threshold, uint8_t count, uint8_t nms,
                            fdf_point *out, size_t cap, uint64_t *offsets) {
    if (!ctx) return FDF_ERR_INVALID_ARGUMENT;
    fdf_status st = check_config(ctx, count, nms);
    if (st != FDF_OK) return st;
    if (!offsets || (!out && cap > 0)) return fail(ctx, FDF_ERR_INVALID_ARGUMENT, "null output pointer");
    if (n_frames == 0 || w < 7 || h < 7) {
        memset(offsets, 0, ((size_t)n_frames + 1) * sizeof(uint64_t));
        return FDF_OK;
    }
    if (!frames) return fail(ctx, FDF_ERR_INVALID_ARGUMENT, "null frame pointer");
    if (pitch < w) return fail(ctx, FDF_ERR_INVALID_ARGUMENT, "pitch %u < width %u", pitch, w);
    if (n_frames == 1) frame_stride = (uint64_t)pitch * h;
    if (frame_stride < (uint64_t)pitch * h) return fail(ctx, FDF_ERR_INVALID_ARGUMENT, "frame_stride smaller than one frame");
    FDF_CUDA(ctx, cudaSetDevice(ctx->device));

    // Frames are staged with a 16-byte multiple pitch (TMA requirement) and processed in sub-batches: the
    // host->device copy of sub-batch j+1 (copy stream) overlaps the kernels of sub-batch j (compute stream) and the
    // device->host copy of its points (back stream).  The packed output needs every sub-batch's base offset, so the
    // host reads each sub-batch's (tiny) offsets before it launches the next one; the copy stream never waits.
    const uint32_t dpitch = (w + 15u) & ~15u;
    const uint64_t dstride = (uint64_t)dpitch * h;
    FDF_CUDA(ctx, ctx->staged_frames.reserve((size_t)dstride * n_frames));
    const size_t worst = (size_t)n_frames * (size_t)(w - 6) * (size_t)(h - 6);
    const size_t dcap = cap < worst ? cap : worst;
    FDF_CUDA(ctx, ctx->staged_points.reserve(dcap ? dcap : 1));
    FDF_CUDA(ctx, ctx->staged_offsets.reserve((size_t)n_frames + 2));
    if (ctx->pinned_offsets_count < (size_t)n_frames + 2) {
        if (ctx->pinned_offsets) cudaFreeHost(ctx->pinned_offsets);
        ctx->pinned_offsets = nullptr;
        ctx->pinned_offsets_count = 0;
        FDF_CUDA(ctx, cudaMallocHost(reinterpret_cast<void **>(&ctx->pinned_offsets),
                                     ((size_t)n_frames + 2) * sizeof(unsigned long long)));
        ctx->pinned_offsets_count = (size_t)n_frames + 2;
    }
    // sub-batch size: about 128 MB of pixels, at least enough strips to fill the GPU several times over
    const unsigned long long sub_bytes = ctx->sub_batch_bytes;
    uint32_t sub = (uint32_t)(sub_bytes / (dstride ? dstride : 1));
    if (sub < 1u) sub = 1u;
    if (sub > n_frames) sub = n_frames;
    const uint32_t nsub = (n_frames + sub - 1) / sub;
    while (ctx->pipe_events.size() < 2 * (size_t)nsub) {
        cudaEvent_t e;
        FDF_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->pipe_events.push_back(e);
    }
    for (uint32_t j = 0; j < nsub; j++) {  // all host->device copies are queued up front
        const uint32_t f0 = j * sub, fn = (f0 + sub <= n_frames) ? sub : n_frames - f0;
        uint8_t *dst = ctx->staged_frames.ptr + (size_t)f0 * dstride;
        const uint8_t *src = frames + (size_t)f0 * frame_stride;
        if (pitch == dpitch && frame_stride == dstride) {
            FDF_CUDA(ctx, cudaMemcpyAsync(dst, src, (size_t)dstride * fn, cudaMemcpyHostToDevice, ctx->copy_stream));
        } else if (frame_stride == (uint64_t)pitch * h) {
            FDF_CUDA(ctx, cudaMemcpy2DAsync(dst, dpitch, src, pitch, w, (size_t)h * fn, cudaMemcpyHostToDevice,
                                            ctx->copy_stream));
        } else {
            for (uint32_t f = 0; f < fn; f++)
                FDF_CUDA(ctx, cudaMemcpy2DAsync(dst + (size_t)f * dstride, dpitch, src + (size_t)f * frame_stride, pitch,
                                                w, h, cudaMemcpyHostToDevice, ctx->copy_stream));
        }
        FDF_CUDA(ctx, cudaEventRecord(ctx->pipe_events[2 * j], ctx->copy_stream));
    }
    uint64_t found = 0;  // keypoints of the sub-batches so far (also beyond the capacity)
    uint32_t flags_all = 0;
    offsets[0] = 0;
    for (uint32_t j = 0; j < nsub; j++) {
        const uint32_t f0 = j * sub, fn = (f0 + sub <= n_frames) ? sub : n_frames - f0;
        const size_t used = found < dcap ? (size_t)found : dcap;
        FDF_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->pipe_events[2 * j], 0));
        unsigned long long *d_off = ctx->staged_offsets.ptr;  // (fn + 1 local offsets; reused by every sub-batch)
        st = fdf_detect_device(ctx, ctx->staged_frames.ptr + (size_t)f0 * dstride, fn, w, h, dpitch, dstride, threshold,
                               count, nms, ctx->staged_points.ptr + used, dcap - used,
                               reinterpret_cast<uint64_t *>(d_off), ctx->stream);
        if (st != FDF_OK) return st;
        FDF_CUDA(ctx, cudaMemcpyAsync(ctx->pinned_offsets, d_off, ((size_t)fn + 1) * sizeof(unsigned long long),
                                      cudaMemcpyDeviceToHost, ctx->stream));
        uint32_t *flags_dev = reinterpret_cast<uint32_t *>(ctx->workspace.ptr + 4);
        FDF_CUDA(ctx, cudaMemcpyAsync(ctx->pinned_offsets + fn + 1, flags_dev, sizeof(uint32_t), cudaMemcpyDeviceToHost,
                                      ctx->stream));
        FDF_CUDA(ctx, cudaEventRecord(ctx->pipe_events[2 * j + 1], ctx->stream));
        FDF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        flags_all |= *reinterpret_cast<const uint32_t *>(ctx->pinned_offsets + fn + 1);
        for (uint32_t f = 0; f < fn; f++) offsets[f0 + f + 1] = found + ctx->pinned_offsets[f + 1];
        const uint64_t got = ctx->pinned_offsets[fn];
        const size_t ncopy = (found + got <= dcap) ? (size_t)got : (dcap - used);
        if (ncopy) {  // this sub-batch's points go home while the next one is computed
            FDF_CUDA(ctx, cudaStreamWaitEvent(ctx->back_stream, ctx->pipe_events[2 * j + 1], 0));
            FDF_CUDA(ctx, cudaMemcpyAsync(out + used, ctx->staged_points.ptr + used, ncopy * sizeof(fdf_point),
                                          cudaMemcpyDeviceToHost, ctx->back_stream));
        }
        found += got;
    }
    FDF_CUDA(ctx, cudaStreamSynchronize(ctx->back_stream));
    if (flags_all & ~4u) return fail(ctx, FDF_ERR_INTERNAL, "device flags 0x%x (look-back or pipeline wait timed out)", flags_all);
    if (found > cap) return fail(ctx, FDF_ERR_CAPACITY, "%llu keypoints found, capacity %zu", (unsigned long long)found, cap);
    if (flags_all) return fail(ctx, FDF_ERR_INTERNAL, "device flags 0x%x (staging buffer overflow)", flags_all);
    return FDF_OK;
}

namespace {
constexpr size_t kSingleHead = 64;          // bytes in front of the points in the single-image device block
constexpr size_t kSingleFirstPoints = 32768;  // points that come back with the first (usually only) device->host copy

fdf_status reserve_pinned(fdf_ctx *ctx, uint8_t **buf, size_t *have, size_t need) {
    if (*have >= need) return FDF_OK;
    if (*buf) cudaFreeHost(*buf);
    *buf = nullptr;
    *have = 0;
    FDF_CUDA(ctx, cudaMallocHost(reinterpret_cast<void **>(buf), need));
    *have = need;
    return FDF_OK;
}
}  // namespace

// One image, one call, as few host round trips as the problem allows: the image goes up with one asynchronous copy
// when the caller's buffer is pinned (cudaHostRegister / cudaHostAlloc), otherwise in four slices through a pinned
// staging buffer of the context (the host copy of slice k + 1 overlaps the DMA of slice k: a pageable cudaMemcpy would
// do the same staging inside the driver, serially); the three kernels follow; offsets, flags and the first 32 K
// points come back in ONE copy from one device block, then one synchronisation.  (fdf_detect_batch is the throughput
// path; this is the latency path of lib.rs:62-64.)
fdf_status fdf_detect(fdf_ctx *ctx, const uint8_t *img, uint32_t w, uint32_t h, uint32_t pitch, uint8_t threshold,
                      uint8_t count, uint8_t nms, fdf_point *out, size_t cap, size_t *n_out) {
    if (!ctx) return FDF_ERR_INVALID_ARGUMENT;
    if (!n_out) return fail(ctx, FDF_ERR_INVALID_ARGUMENT, "null n_out");
    *n_out = 0;
    fdf_status st = check_config(ctx, count, nms);
    if (st != FDF_OK) return st;
    if (!out && cap > 0) return fail(ctx, FDF_ERR_INVALID_ARGUMENT, "null output pointer");
    if (w < 7 || h < 7) return FDF_OK;  // nothing can be a keypoint
    if (!img) return fail(ctx, FDF_ERR_INVALID_ARGUMENT, "null image pointer");
    if (pitch < w) return fail(ctx, FDF_ERR_INVALID_ARGUMENT, "pitch %u < width %u", pitch, w);
    FDF_CUDA(ctx, cudaSetDevice(ctx->device));

    const uint32_t dpitch = (w + 15u) & ~15u;
    const size_t worst = (size_t)(w - 6) * (size_t)(h - 6), dcap = cap < worst ? cap : worst;
    // points that come back with the first (usually only) copy: 1.5 x the previous call's count + 1024, at least 4096
    size_t first = ctx->single_hint + ctx->single_hint / 2 + 1024;
    if (first < 4096) first = 4096;
    if (first > kSingleFirstPoints) first = kSingleFirstPoints;
    if (first > dcap) first = dcap;
    FDF_CUDA(ctx, ctx->staged_frames.reserve((size_t)dpitch * h));
    FDF_CUDA(ctx, ctx->single_block.reserve(kSingleHead + (dcap ? dcap : 1) * sizeof(fdf_point)));
    if ((st = reserve_pinned(ctx, &ctx->pinned_head, &ctx->pinned_head_bytes, kSingleHead + kSingleFirstPoints * sizeof(fdf_point))) != FDF_OK)
        return st;
    uint8_t *block = ctx->single_block.ptr;
    uint64_t *d_offsets = reinterpret_cast<uint64_t *>(block);
    fdf_point *d_points = reinterpret_cast<fdf_point *>(block + kSingleHead);

    // ---- the image ----
    cudaPointerAttributes attr;
    const bool pinned = cudaPointerGetAttributes(&attr, img) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();  // (an unregistered pointer is not an error here)
    if (pinned) {
        FDF_CUDA(ctx, cudaMemcpy2DAsync(ctx->staged_frames.ptr, dpitch, img, pitch, w, h, cudaMemcpyHostToDevice, ctx->stream));
    } else {
        if ((st = reserve_pinned(ctx, &ctx->pinned_in, &ctx->pinned_in_bytes, (size_t)dpitch * h)) != FDF_OK) return st;
        const uint32_t slices = h >= 64 ? 4u : 1u;
        for (uint32_t k = 0; k < slices; k++) {
            const uint32_t r0 = (uint32_t)((uint64_t)h * k / slices), r1 = (uint32_t)((uint64_t)h * (k + 1) / slices);
            uint8_t *stage = ctx->pinned_in + (size_t)r0 * dpitch;
            if (pitch == dpitch) {
                memcpy(stage, img + (size_t)r0 * pitch, (size_t)(r1 - r0) * pitch);
            } else {
                for (uint32_t r = r0; r < r1; r++) memcpy(stage + (size_t)(r - r0) * dpitch, img + (size_t)r * pitch, w);
            }
            FDF_CUDA(ctx, cudaMemcpyAsync(ctx->staged_frames.ptr + (size_t)r0 * dpitch, stage, (size_t)(r1 - r0) * dpitch,
                                          cudaMemcpyHostToDevice, ctx->stream));
        }
    }
    // ---- the kernels ----
    ctx->flags_copy_next = reinterpret_cast<uint32_t *>(block + 16);  // (the gather kernel puts the launch's flags there)
    st = fdf_detect_device(ctx, ctx->staged_frames.ptr, 1, w, h, dpitch, (uint64_t)dpitch * h, threshold, count, nms,
                           d_points, dcap, d_offsets, ctx->stream);
    ctx->flags_copy_next = nullptr;
    if (st != FDF_OK) return st;
    // ---- offsets + flags + the first points: one copy, one synchronisation ----
    FDF_CUDA(ctx, cudaMemcpyAsync(ctx->pinned_head, block, kSingleHead + first * sizeof(fdf_point), cudaMemcpyDeviceToHost,
                                  ctx->stream));
    FDF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const uint64_t found = reinterpret_cast<const uint64_t *>(ctx->pinned_head)[1];
    const uint32_t flags = *reinterpret_cast<const uint32_t *>(ctx->pinned_head + 16);
    *n_out = (size_t)found;
    ctx->single_hint = (size_t)found;
    if (flags & ~4u) return fail(ctx, FDF_ERR_INTERNAL, "device flags 0x%x (look-back or pipeline wait timed out)", flags);
    if (found > cap) return fail(ctx, FDF_ERR_CAPACITY, "%llu keypoints found, capacity %zu", (unsigned long long)found, cap);
    if (flags) return fail(ctx, FDF_ERR_INTERNAL, "device flags 0x%x (staging buffer overflow)", flags);
    const size_t n = (size_t)found;
    memcpy(out, ctx->pinned_head + kSingleHead, (n < first ? n : first) * sizeof(fdf_point));
    if (n > first)  // (rare: more than 32 K keypoints in one image)
        FDF_CUDA(ctx, cudaMemcpy(out + first, d_points + first, (n - first) * sizeof(fdf_point), cudaMemcpyDeviceToHost));
    return FDF_OK;
}

// ---- streaming form of fdf_detect ------------------------------------------------------------------
struct fdf_pipe {
    struct Slot {
        uint8_t *pinned_in = nullptr;   // staging for pageable images
        uint8_t *pinned_out = nullptr;  // [offsets u64 x 2 | flags @16 | pad to 64 | points]
        uint8_t *d_frame = nullptr;
        uint8_t *d_block = nullptr;     // same layout as pinned_out
        cudaEvent_t landed = nullptr, kernels = nullptr, done = nullptr;
        size_t first = 0;               // points that come back with the asynchronous copy
        bool empty = false;             // image smaller than 7 x 7: no device work
    };
    fdf_ctx *ctx = nullptr;
    uint32_t depth = 0, max_w = 0, max_h = 0, head = 0, tail = 0, in_flight = 0;
    size_t cap = 0, hint = 0;
    std::vector<Slot> slots;
};

fdf_status fdf_pipe_create(fdf_ctx *ctx, uint32_t depth, uint32_t max_w, uint32_t max_h, size_t cap, fdf_pipe **out_pipe) {
    if (!ctx || !out_pipe) return FDF_ERR_INVALID_ARGUMENT;
    *out_pipe = nullptr;
    if (depth == 0 || depth > 64 || max_w == 0 || max_h == 0)
        return fail(ctx, FDF_ERR_INVALID_ARGUMENT, "pipe needs 1..64 slots and a non-empty maximum image size");
    FDF_CUDA(ctx, cudaSetDevice(ctx->device));
    fdf_pipe *pipe = new (std::nothrow) fdf_pipe();
    if (!pipe) return fail(ctx, FDF_ERR_INTERNAL, "out of host memory");
    pipe->ctx = ctx, pipe->depth = depth, pipe->max_w = max_w, pipe->max_h = max_h, pipe->cap = cap;
    pipe->slots.resize(depth);
    const size_t frame_bytes = (size_t)((max_w + 15u) & ~15u) * max_h;
    const size_t block_bytes = kSingleHead + (cap ? cap : 1) * sizeof(fdf_point);
    for (auto &s : pipe->slots) {
        cudaError_t e = cudaMallocHost(reinterpret_cast<void **>(&s.pinned_in), frame_bytes);
        if (e == cudaSuccess) e = cudaMallocHost(reinterpret_cast<void **>(&s.pinned_out), block_bytes);
        if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void **>(&s.d_frame), frame_bytes);
        if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void **>(&s.d_block), block_bytes);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s.landed, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s.kernels, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming);
        if (e != cudaSuccess) {
            fdf_pipe_destroy(pipe);
            return fail(ctx, FDF_ERR_CUDA, "pipe allocation failed: %s", cudaGetErrorString(e));
        }
    }
    *out_pipe = pipe;
    return FDF_OK;
}

void fdf_pipe_destroy(fdf_pipe *pipe) {
    if (!pipe) return;
    cudaSetDevice(pipe->ctx->device);
    cudaStreamSynchronize(pipe->ctx->copy_stream);
    cudaStreamSynchronize(pipe->ctx->stream);
    cudaStreamSynchronize(pipe->ctx->back_stream);
    for (auto &s : pipe->slots) {
        if (s.pinned_in) cudaFreeHost(s.pinned_in);
        if (s.pinned_out) cudaFreeHost(s.pinned_out);
        if (s.d_frame) cudaFree(s.d_frame);
        if (s.d_block) cudaFree(s.d_block);
        if (s.landed) cudaEventDestroy(s.landed);
        if (s.kernels) cudaEventDestroy(s.kernels);
        if (s.done) cudaEventDestroy(s.done);
    }
    delete pipe;
}

uint32_t fdf_pipe_in_flight(const fdf_pipe *pipe) { return pipe ? pipe->in_flight : 0u; }

fdf_status fdf_pipe_submit(fdf_pipe *pipe, const uint8_t *img, uint32_t w, uint32_t h, uint32_t pitch, uint8_t threshold,
                           uint8_t count, uint8_t nms) {
    if (!pipe) return FDF_ERR_INVALID_ARGUMENT;
    fdf_ctx *ctx = pipe->ctx;
    fdf_status st = check_config(ctx, count, nms);
    if (st != FDF_OK) return st;
    if (pipe->in_flight == pipe->depth) return fail(ctx, FDF_ERR_BUSY, "%u images in flight: collect one first", pipe->depth);
    if (w > pipe->max_w || h > pipe->max_h)
        return fail(ctx, FDF_ERR_INVALID_ARGUMENT, "image %u x %u exceeds the pipe's %u x %u", w, h, pipe->max_w, pipe->max_h);
    fdf_pipe::Slot &s = pipe->slots[pipe->tail];
    s.empty = w < 7 || h < 7;  // nothing can be a keypoint
    if (!s.empty) {
        if (!img) return fail(ctx, FDF_ERR_INVALID_ARGUMENT, "null image pointer");
        if (pitch < w) return fail(ctx, FDF_ERR_INVALID_ARGUMENT, "pitch %u < width %u", pitch, w);
        FDF_CUDA(ctx, cudaSetDevice(ctx->device));
        const uint32_t dpitch = (w + 15u) & ~15u;
        cudaPointerAttributes attr;
        const bool pinned = cudaPointerGetAttributes(&attr, img) == cudaSuccess && attr.type == cudaMemoryTypeHost;
        cudaGetLastError();  // (an unregistered pointer is not an error here)
        if (pinned) {
            FDF_CUDA(ctx, cudaMemcpy2DAsync(s.d_frame, dpitch, img, pitch, w, h, cudaMemcpyHostToDevice, ctx->copy_stream));
        } else {
            if (pitch == dpitch) {
                memcpy(s.pinned_in, img, (size_t)(h - 1) * pitch + w);
            } else {
                for (uint32_t r = 0; r < h; r++) memcpy(s.pinned_in + (size_t)r * dpitch, img + (size_t)r * pitch, w);
            }
            FDF_CUDA(ctx, cudaMemcpyAsync(s.d_frame, s.pinned_in, (size_t)dpitch * h, cudaMemcpyHostToDevice, ctx->copy_stream));
        }
        FDF_CUDA(ctx, cudaEventRecord(s.landed, ctx->copy_stream));
        FDF_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, s.landed, 0));
        const size_t worst = (size_t)(w - 6) * (size_t)(h - 6), dcap = pipe->cap < worst ? pipe->cap : worst;
        ctx->flags_copy_next = reinterpret_cast<uint32_t *>(s.d_block + 16);
        st = fdf_detect_device(ctx, s.d_frame, 1, w, h, dpitch, (uint64_t)dpitch * h, threshold, count, nms,
                               reinterpret_cast<fdf_point *>(s.d_block + kSingleHead), dcap,
                               reinterpret_cast<uint64_t *>(s.d_block), ctx->stream);
        ctx->flags_copy_next = nullptr;
        if (st != FDF_OK) return st;
        FDF_CUDA(ctx, cudaEventRecord(s.kernels, ctx->stream));
        FDF_CUDA(ctx, cudaStreamWaitEvent(ctx->back_stream, s.kernels, 0));
        s.first = pipe->hint + pipe->hint / 2 + 1024;
        if (s.first < 4096) s.first = 4096;
        if (s.first > dcap) s.first = dcap;
        FDF_CUDA(ctx, cudaMemcpyAsync(s.pinned_out, s.d_block, kSingleHead + s.first * sizeof(fdf_point),
                                      cudaMemcpyDeviceToHost, ctx->back_stream));
        FDF_CUDA(ctx, cudaEventRecord(s.done, ctx->back_stream));
    }
    pipe->tail = (pipe->tail + 1u) % pipe->depth;
    pipe->in_flight++;
    return FDF_OK;
}

fdf_status fdf_pipe_collect(fdf_pipe *pipe, fdf_point *out, size_t cap, size_t *n_out) {
    if (!pipe) return FDF_ERR_INVALID_ARGUMENT;
    fdf_ctx *ctx = pipe->ctx;
    if (!n_out) return fail(ctx, FDF_ERR_INVALID_ARGUMENT, "null n_out");
    *n_out = 0;
    if (pipe->in_flight == 0) return fail(ctx, FDF_ERR_INVALID_ARGUMENT, "no image in flight");
    if (!out && cap > 0) return fail(ctx, FDF_ERR_INVALID_ARGUMENT, "null output pointer");
    fdf_pipe::Slot &s = pipe->slots[pipe->head];
    pipe->head = (pipe->head + 1u) % pipe->depth;  // (the image is retired whatever happens below)
    pipe->in_flight--;
    if (s.empty) return FDF_OK;
    FDF_CUDA(ctx, cudaSetDevice(ctx->device));
    FDF_CUDA(ctx, cudaEventSynchronize(s.done));
    const uint64_t found = reinterpret_cast<const uint64_t *>(s.pinned_out)[1];
    const uint32_t flags = *reinterpret_cast<const uint32_t *>(s.pinned_out + 16);
    *n_out = (size_t)found;
    pipe->hint = (size_t)found;
    if (flags & ~4u) return fail(ctx, FDF_ERR_INTERNAL, "device flags 0x%x (look-back or pipeline wait timed out)", flags);
    if (found > pipe->cap) return fail(ctx, FDF_ERR_CAPACITY, "%llu keypoints found, the pipe was created for %zu", (unsigned long long)found, pipe->cap);
    if (found > cap) return fail(ctx, FDF_ERR_CAPACITY, "%llu keypoints found, capacity %zu", (unsigned long long)found, cap);
    if (flags) return fail(ctx, FDF_ERR_INTERNAL, "device flags 0x%x (staging buffer overflow)", flags);
    const size_t n = (size_t)found;
    memcpy(out, s.pinned_out + kSingleHead, (n < s.first ? n : s.first) * sizeof(fdf_point));
    if (n > s.first)  // (more than the asynchronous copy brought back: fetch the rest now)
        FDF_CUDA(ctx, cudaMemcpy(out + s.first, reinterpret_cast<const fdf_point *>(s.d_block + kSingleHead) + s.first,
                                 (n - s.first) * sizeof(fdf_point), cudaMemcpyDeviceToHost));
    return FDF_OK;
}

namespace {
fdf_status rgb_to_grey(fdf_ctx *ctx, const uint8_t *d_rgb, uint32_t n_frames, uint32_t w, uint32_t h, uint32_t rgb_pitch,
                       uint64_t rgb_frame_stride, uint8_t *d_luma, uint32_t luma_pitch, uint64_t luma_frame_stride,
                       int kind, void *stream_handle) {
    if (!ctx) return FDF_ERR_INVALID_ARGUMENT;
    if (n_frames == 0 || w == 0 || h == 0) return FDF_OK;
    if (!d_rgb || !d_luma) return fail(ctx, FDF_ERR_INVALID_ARGUMENT, "null image pointer");
    if ((uint64_t)rgb_pitch < 3ull * w || luma_pitch < w)
        return fail(ctx, FDF_ERR_INVALID_ARGUMENT, "pitch smaller than one row (rgb %u, luma %u, width %u)", rgb_pitch,
                    luma_pitch, w);
    if (n_frames > 1 && (rgb_frame_stride < (uint64_t)rgb_pitch * h || luma_frame_stride < (uint64_t)luma_pitch * h))
        return fail(ctx, FDF_ERR_INVALID_ARGUMENT, "frame stride smaller than one frame");
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_handle);  // NULL = the default stream
    FDF_CUDA(ctx, cudaSetDevice(ctx->device));
    FDF_CUDA(ctx, fdf::launch_luma(d_rgb, n_frames, w, h, rgb_pitch, rgb_frame_stride, d_luma, luma_pitch,
                                   luma_frame_stride, kind, stream));
    ctx->launches += 1;
    return FDF_OK;
}
}  // namespace

fdf_status fdf_rgb8_to_luma8_device(fdf_ctx *ctx, const uint8_t *d_rgb, uint32_t n_frames, uint32_t w, uint32_t h,
                                    uint32_t rgb_pitch, uint64_t rgb_frame_stride, uint8_t *d_luma,
                                    uint32_t luma_pitch, uint64_t luma_frame_stride, void *stream_handle) {
    return rgb_to_grey(ctx, d_rgb, n_frames, w, h, rgb_pitch, rgb_frame_stride, d_luma, luma_pitch, luma_frame_stride, 0,
                       stream_handle);
}

fdf_status fdf_rgb8_to_grey_sum3_device(fdf_ctx *ctx, const uint8_t *d_rgb, uint32_t n_frames, uint32_t w, uint32_t h,
                                        uint32_t rgb_pitch, uint64_t rgb_frame_stride, uint8_t *d_grey,
                                        uint32_t grey_pitch, uint64_t grey_frame_stride, void *stream_handle) {
    return rgb_to_grey(ctx, d_rgb, n_frames, w, h, rgb_pitch, rgb_frame_stride, d_grey, grey_pitch, grey_frame_stride, 1,
                       stream_handle);
}

fdf_status fdf_detect_rgb8(fdf_ctx *ctx, const uint8_t *rgb, uint32_t w, uint32_t h, uint32_t rgb_pitch,
                           uint8_t threshold, uint8_t count, uint8_t nms, fdf_point *out, size_t cap,
                           size_t *n_out) {
    if (!ctx) return FDF_ERR_INVALID_ARGUMENT;
    if (!n_out) return fail(ctx, FDF_ERR_INVALID_ARGUMENT, "null n_out");
    *n_out = 0;
    fdf_status st = check_config(ctx, count, nms);
    if (st != FDF_OK) return st;
    if (!out && cap > 0) return fail(ctx, FDF_ERR_INVALID_ARGUMENT, "null output pointer");
    if (w < 7 || h < 7) return FDF_OK;
    if (!rgb) return fail(ctx, FDF_ERR_INVALID_ARGUMENT, "null image pointer");
    if ((uint64_t)rgb_pitch < 3ull * w) return fail(ctx, FDF_ERR_INVALID_ARGUMENT, "rgb_pitch %u < 3 * width", rgb_pitch);
    FDF_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint32_t dpitch = (w + 15u) & ~15u, rpitch = (3u * w + 3u) & ~3u;
    const size_t worst = (size_t)(w - 6) * (size_t)(h - 6), dcap = cap < worst ? cap : worst;
    FDF_CUDA(ctx, ctx->staged_rgb.reserve((size_t)rpitch * h));
    FDF_CUDA(ctx, ctx->staged_frames.reserve((size_t)dpitch * h));
    FDF_CUDA(ctx, ctx->staged_points.reserve(dcap ? dcap : 1));
    FDF_CUDA(ctx, ctx->staged_offsets.reserve(4));
    FDF_CUDA(ctx, cudaMemcpy2DAsync(ctx->staged_rgb.ptr, rpitch, rgb, rgb_pitch, 3u * (size_t)w, h, cudaMemcpyHostToDevice,
                                    ctx->stream));
    st = fdf_rgb8_to_luma8_device(ctx, ctx->staged_rgb.ptr, 1, w, h, rpitch, (uint64_t)rpitch * h, ctx->staged_frames.ptr,
                                  dpitch, (uint64_t)dpitch * h, ctx->stream);
    if (st != FDF_OK) return st;
    st = fdf_detect_device(ctx, ctx->staged_frames.ptr, 1, w, h, dpitch, (uint64_t)dpitch * h, threshold, count, nms,
                           ctx->staged_points.ptr, dcap, reinterpret_cast<uint64_t *>(ctx->staged_offsets.ptr), ctx->stream);
    if (st != FDF_OK) return st;
    unsigned long long offs[2] = {0, 0};
    uint32_t flags = 0;
    FDF_CUDA(ctx, cudaMemcpyAsync(offs, ctx->staged_offsets.ptr, sizeof(offs), cudaMemcpyDeviceToHost, ctx->stream));
    FDF_CUDA(ctx, cudaMemcpyAsync(&flags, ctx->workspace.ptr + 4, sizeof(flags), cudaMemcpyDeviceToHost, ctx->stream));
    FDF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *n_out = (size_t)offs[1];
    if (flags & ~4u) return fail(ctx, FDF_ERR_INTERNAL, "device flags 0x%x", flags);
    if (offs[1] > cap) return fail(ctx, FDF_ERR_CAPACITY, "%llu keypoints found, capacity %zu", offs[1], cap);
    if (flags) return fail(ctx, FDF_ERR_INTERNAL, "device flags 0x%x (staging buffer overflow)", flags);
    const size_t ncopy = offs[1] < dcap ? (size_t)offs[1] : dcap;
    if (ncopy) FDF_CUDA(ctx, cudaMemcpy(out, ctx->staged_points.ptr, ncopy * sizeof(fdf_point), cudaMemcpyDeviceToHost));
    return FDF_OK;
}

#ifdef FDF_CHECKS
// (checks builds only: tools/build_variant.sh checks -DFDF_CHECKS) source line of a failed shared-memory index check since the last call
fdf_status fdf_debug_check_failure(fdf_ctx *ctx, int *line) {
    if (!ctx || !line) return FDF_ERR_INVALID_ARGUMENT;
    FDF_CUDA(ctx, cudaDeviceSynchronize());
    FDF_CUDA(ctx, fdf::read_check_failure(line));
    return FDF_OK;
}
#endif

#ifdef FDF_TRACE
// (debug builds only) the timeline table of the last launch: [cta 4][warp 16][chunk 200][slot 12] clock64 values
fdf_status fdf_debug_trace(fdf_ctx *ctx, int64_t *out, size_t bytes) {
    if (!ctx || !out) return FDF_ERR_INVALID_ARGUMENT;
    FDF_CUDA(ctx, cudaDeviceSynchronize());
    FDF_CUDA(ctx, fdf::read_trace(reinterpret_cast<long long *>(out), bytes));
    return FDF_OK;
}
#endif

fdf_status fdf_synth_frames_device(fdf_ctx *ctx, uint8_t *d_frames, uint32_t n_frames, uint32_t w, uint32_t h,
                                   uint32_t pitch, uint64_t frame_stride, uint64_t seed, uint32_t first_frame,
                                   uint32_t kind, uint32_t amp, void *stream_handle) {
    if (!ctx || !d_frames) return FDF_ERR_INVALID_ARGUMENT;
    if (pitch < w || kind > 1) return fail(ctx, FDF_ERR_INVALID_ARGUMENT, "bad synth arguments");
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_handle);  // NULL = the default stream
    FDF_CUDA(ctx, cudaSetDevice(ctx->device));
    FDF_CUDA(ctx, fdf::launch_synth(d_frames, n_frames, w, h, pitch, frame_stride, seed, first_frame, kind, amp, stream));
    ctx->launches += 1;
    return FDF_OK;
}

}  // extern "C"
