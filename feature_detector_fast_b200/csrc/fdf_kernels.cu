// fdf_kernels.cu -- sm_100a kernels of the FAST-n detection path.
//
// One persistent kernel does the whole path for a batch of frames (replaces fast_simd.rs:301-620):
//
//   work item  = (frame, strip of full-width rows), handed out through an atomic ticket in row-major
//                order, which is what makes the decoupled look-back deadlock-free.
//   8 compute warps, per chunk of a strip:
//     TMA 3-D tiled load (u8 tile 256 x (SR+6), zero-filled outside the image, double buffered,
//     completion on an mbarrier)                                                -> shared memory
//     phase A  : dense SWAR filter, 16 centres per thread (LDS.128 + PRMT + VABSDIFF4 + LOP3),
//                survivors pushed to the warp's own candidate queue            (fast_simd.rs:368-520)
//     phase B  : one lane per candidate: 16 ring bytes -> brighter/darker 16-bit masks ->
//                rotate-AND arc test -> score in registers -> tagged score plane (fast_simd.rs:115-297,
//                                                                                623-749)
//     one named barrier, then
//     NMS pass : strict 3x3 maximum on the shared-memory score plane          (fast_simd.rs:588-616)
//                survivors set one bit in the strip's shared-memory bit plane (double buffered)
//   1 emit warp, per finished strip (while the compute warps already work on the next one):
//     popcount of the bit plane, decoupled look-back over all earlier items of the whole batch,
//     then the bits are expanded to (x, y) points at their final position: output is packed and
//     row-major per frame (fast_simd.rs:550, 596-613).
#include "fdf_kernels.cuh"

#include "fdf_core.cuh"
#include "fdf_strip.cuh"
#include "fdf_synth.cuh"

namespace fdf {
namespace {

constexpr uint32_t kFlagLookbackTimeout = 1u;
constexpr uint32_t kFlagTmaTimeout = 2u;
constexpr unsigned long long kStatusAggregate = 1ull << 62;
constexpr unsigned long long kStatusPrefix = 2ull << 62;
constexpr unsigned long long kStatusValueMask = (1ull << 62) - 1ull;
constexpr uint32_t kSpinLimit = 1u << 22;

// ---- PTX wrappers ------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// barrier among the compute warps only (the emit warp never joins it)
__device__ __forceinline__ void bar_compute() {
    asm volatile("bar.sync 1, %0;" ::"n"(kComputeThreads) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity, uint32_t *flags) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > kSpinLimit) {  // never expected; turns a would-be hang into an error flag
            atomicOr(flags, kFlagTmaTimeout);
            break;
        }
    }
}

// TMA: 3-D tiled load global -> shared, completion signalled on an mbarrier (SASS: UTMALDG)
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *tmap, int x, int y, int z, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
        : "memory");
}

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}

__device__ __forceinline__ unsigned long long ld_relaxed_gpu(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void st_relaxed_gpu(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// ---- shared-memory carve-up ------------------------------------------------------------------
template <int MODE, int SR>
struct Layout {
    static constexpr int TR = tile_rows(SR);
    static constexpr int tile_bytes = TR * kTileW;  // one TMA box
    static constexpr int plane_off = 2 * tile_bytes;
    static constexpr int plane_bytes = (MODE == NMS_OFF) ? 0 : SR * kPlaneW * 2;  // u16: tag << 12 | score
    static constexpr int queue_off = plane_off + plane_bytes;
    static constexpr int queue_per_warp = (SR / kComputeWarps) * kTileW;  // entries: every pixel of the warp's rows
    static constexpr int queue_bytes = kComputeWarps * queue_per_warp * 2;
    static constexpr int misc_off = queue_off + queue_bytes;
    static constexpr int misc_bytes = 128;
    static constexpr int bits_off = misc_off + misc_bytes;  // two bit planes of out_rows x words_per_row words
    static_assert(tile_bytes % 128 == 0, "TMA destination must stay 128-byte aligned");
    static_assert(SR % 16 == 0 && SR <= 64, "phase A walks rows in steps of 16; queue entries hold 6 row bits");
};

// ---- decoupled look-back (one warp) -------------------------------------------------------------
// status[i]: bits 63:62 = 0 empty / 1 aggregate of item i / 2 inclusive prefix up to item i.
__device__ __forceinline__ unsigned long long lookback(unsigned long long *status, uint32_t item, uint32_t total,
                                                       int lane, uint32_t *flags) {
    if (item == 0) {
        if (lane == 0) st_relaxed_gpu(&status[0], kStatusPrefix | total);
        return 0ull;
    }
    if (lane == 0) st_relaxed_gpu(&status[item], kStatusAggregate | total);
    unsigned long long excl = 0ull;
    long long j = (long long)item - 1;
    uint32_t spins = 0;
    while (true) {
        const long long idx = j - lane;  // lane 0 looks at the nearest predecessor
        const unsigned long long s = idx >= 0 ? ld_relaxed_gpu(&status[idx]) : kStatusPrefix;
        const uint32_t flag = (uint32_t)(s >> 62);
        const uint32_t prefix_lanes = __ballot_sync(0xffffffffu, flag == 2u);
        const uint32_t empty_lanes = __ballot_sync(0xffffffffu, flag == 0u);
        // lanes 0 .. (first lane holding a prefix) are the ones whose values are needed
        const uint32_t need = prefix_lanes ? ((2u << (__ffs(prefix_lanes) - 1)) - 1u) : 0xffffffffu;
        if (empty_lanes & need) {
            if (++spins > kSpinLimit) {
                if (lane == 0) atomicOr(flags, kFlagLookbackTimeout);
                break;
            }
            __nanosleep(32);
            continue;
        }
        unsigned long long v = ((need >> lane) & 1u) ? (s & kStatusValueMask) : 0ull;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
        excl += v;
        if (prefix_lanes) break;
        j -= 32;
    }
    if (lane == 0) st_relaxed_gpu(&status[item], kStatusPrefix | (excl + total));
    return excl;
}

// ---- the detection kernel ----------------------------------------------------------------------
template <int MODE, int SR>
__global__ void __launch_bounds__(kThreads, 2)
fdf_detect_kernel(const __grid_constant__ CUtensorMap tmap, const DetectParams p) {
    using L = Layout<MODE, SR>;
    constexpr int HS = (MODE == NMS_OFF) ? 0 : 1;  // score halo (rows and columns) needed by the 3x3 NMS
    constexpr int OUT_R = out_rows(MODE, SR);

    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *tiles = smem;
    uint16_t *plane = reinterpret_cast<uint16_t *>(smem + L::plane_off);
    uint16_t *queue = reinterpret_cast<uint16_t *>(smem + L::queue_off);
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + L::misc_off);             // [2] tile landed
    uint64_t *bits_full = reinterpret_cast<uint64_t *>(smem + L::misc_off + 16);       // [2] strip finished
    uint64_t *bits_empty = reinterpret_cast<uint64_t *>(smem + L::misc_off + 32);      // [2] strip emitted
    uint32_t *qcount = reinterpret_cast<uint32_t *>(smem + L::misc_off + 48);          // [8] per-warp queue fill
    uint32_t *s_item = reinterpret_cast<uint32_t *>(smem + L::misc_off + 80);          // [2] item of each bit plane
    uint32_t *bits = reinterpret_cast<uint32_t *>(smem + L::bits_off);                  // [2][OUT_R][words_per_row]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int W = (int)p.w, H = (int)p.h;
    const int NC = (int)p.chunks_per_strip;
    const int WW = (int)p.words_per_row;
    const int nwords = OUT_R * WW;
    const uint32_t total_items = p.n_frames * p.strips_per_frame;

    if (tid == 0) {
        tma_prefetch_desc(&tmap);
        mbar_init(&full_bar[0], 1);
        mbar_init(&full_bar[1], 1);
        mbar_init(&bits_full[0], kComputeWarps);
        mbar_init(&bits_full[1], kComputeWarps);
        mbar_init(&bits_empty[0], 1);
        mbar_init(&bits_empty[1], 1);
        fence_mbar_init();
    }
    for (int i = tid; i < 2 * nwords; i += kThreads) bits[i] = 0u;
    __syncthreads();

    if (warp < kComputeWarps) {
        // ======================= compute warps =======================
        const int t = (int)p.threshold, n = (int)p.count;
        const uint32_t kbias = filter_kbias(p.threshold);
        uint16_t *wqueue = queue + warp * L::queue_per_warp;
        uint32_t *wcount = &qcount[warp];
        uint32_t gc = 0;  // chunks processed by this CTA so far: tile stage = gc & 1, mbarrier parity = (gc >> 1) & 1
        for (uint32_t it = 0;; it++) {
            const int buf = (int)(it & 1u);
            if (tid == 0) {
                mbar_wait(&bits_empty[buf], ((it >> 1) & 1u) ^ 1u, p.flags);  // bit plane `buf` emitted and zeroed
                s_item[buf] = atomicAdd(p.ticket, 1u);  // items start in scan order => look-back cannot deadlock
            }
            bar_compute();  // also: every warp has finished the previous strip (tiles, plane and queues are free)
            const uint32_t item = s_item[buf];
            if (item >= total_items) {
                if (lane == 0) mbar_arrive(&bits_full[buf]);  // hand the end marker to the emit warp
                break;
            }
            const uint32_t frame = item / p.strips_per_frame;
            const uint32_t strip = item - frame * p.strips_per_frame;
            const int ty0 = first_out_row(MODE) + (int)strip * OUT_R - HS - 3;  // image row of tile row 0
            uint32_t *sbits = bits + buf * nwords;

            if (tid == 0) {
                const int pre = NC < 2 ? NC : 2;
                for (int c = 0; c < pre; c++) {
                    const uint32_t stage = (gc + (uint32_t)c) & 1u;
                    mbar_expect_tx(&full_bar[stage], (uint32_t)L::tile_bytes);
                    tma_load_3d(tiles + stage * L::tile_bytes, &tmap, c * kChunkW - kTileLead, ty0, (int)frame,
                                &full_bar[stage]);
                }
            }
            for (int c = 0; c < NC; c++, gc++) {
                const uint32_t stage = gc & 1u;
                const uint8_t *tile = tiles + stage * L::tile_bytes;
                const ChunkGeo g = make_geo<MODE>(W, H, WW, (int)strip, c, SR);
                const uint32_t tag = (uint32_t)(c % kTagPeriod) + 1u;
                if (MODE != NMS_OFF && tag == 1u) {  // (re)start the tag sequence on a cleared plane
                    if (c != 0) bar_compute();        // every warp is done suppressing the previous chunk
                    uint4 *pz = reinterpret_cast<uint4 *>(plane) + tid;
#pragma unroll
                    for (int i = 0; i < L::plane_bytes / 16 / kComputeThreads; i++)
                        pz[i * kComputeThreads] = make_uint4(0u, 0u, 0u, 0u);
                    bar_compute();
                }
                if (lane == 0) *wcount = 0u;
                mbar_wait(&full_bar[stage], (gc >> 1) & 1u, p.flags);
                __syncwarp();

                phase_a<MODE, SR>(warp, lane, tile, wqueue, wcount, g, kbias);
                __syncwarp();  // the warp's queue is complete
                const uint32_t qn = *wcount;
                phase_b<MODE, SR>(lane, qn, tile, wqueue, plane, sbits, g, t, n, tag);

                bar_compute();  // tile[stage] is free again; every warp's scores of this chunk are in the plane
                if (tid == 0 && c + 2 < NC) {
                    mbar_expect_tx(&full_bar[stage], (uint32_t)L::tile_bytes);
                    tma_load_3d(tiles + stage * L::tile_bytes, &tmap, (c + 2) * kChunkW - kTileLead, ty0, (int)frame,
                                &full_bar[stage]);
                }
                if (MODE != NMS_OFF) nms_pass<MODE, SR>(lane, qn, wqueue, plane, sbits, g, tag);
                __syncwarp();
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bits_full[buf]);  // this warp's bits of the strip are set
        }
    } else {
        // ======================= emit warp =======================
        for (uint32_t it = 0;; it++) {
            const int buf = (int)(it & 1u);
            mbar_wait(&bits_full[buf], (it >> 1) & 1u, p.flags);
            const uint32_t item = s_item[buf];
            if (item >= total_items) break;
            const uint32_t frame = item / p.strips_per_frame;
            const uint32_t strip = item - frame * p.strips_per_frame;
            const int y0 = first_out_row(MODE) + (int)strip * OUT_R;  // first row this strip emits
            uint32_t *sbits = bits + buf * nwords;

            uint32_t cnt = 0;
            for (int i = lane; i < nwords; i += 32) cnt += (uint32_t)__popc(sbits[i]);
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, d);  // strip total
            unsigned long long o = lookback(p.status, item, cnt, lane, p.flags);
            if (lane == 0) {
                if (item == 0) p.offsets[0] = 0ull;
                if (strip == p.strips_per_frame - 1) p.offsets[frame + 1] = o + cnt;
            }
            if (cnt != 0u) {
                int i = lane;
                int row = i / WW, col = i - row * WW;
                for (int base = 0; base < nwords; base += 32, i += 32) {
                    const uint32_t m = i < nwords ? sbits[i] : 0u;
                    const uint32_t c = (uint32_t)__popc(m);
                    uint32_t incl = c;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
                        if (lane >= d) incl += v;
                    }
                    if (m != 0u) {
                        emit_word(m, (uint32_t)col * 32u, (uint32_t)(y0 + row), o + (incl - c), p.cap, p.out);
                        sbits[i] = 0u;  // leave the plane zeroed for the strip after next
                    }
                    o += __shfl_sync(0xffffffffu, incl, 31);
                    col += 32;
                    while (col >= WW) {
                        col -= WW;
                        row++;
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bits_empty[buf]);
        }
    }
}

__global__ void fdf_synth_kernel(uint8_t *frames, uint32_t n_frames, uint32_t w, uint32_t h, uint32_t pitch,
                                 unsigned long long frame_stride, unsigned long long seed, uint32_t first_frame,
                                 uint32_t kind, uint32_t amp) {
    const unsigned long long w4 = (w + 3u) / 4u;
    const unsigned long long total = (unsigned long long)n_frames * h * w4;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const uint32_t x4 = (uint32_t)(i % w4);
        const uint32_t y = (uint32_t)((i / w4) % h);
        const uint32_t f = (uint32_t)(i / (w4 * h));
        const uint64_t key = synth_frame_key(seed, first_frame + f);
        uint8_t *row = frames + (size_t)f * frame_stride + (size_t)y * pitch;
        for (uint32_t b = 0; b < 4u; b++) {
            const uint32_t x = x4 * 4u + b;
            if (x < w) row[x] = synth_pixel(key, x, y, kind, amp);
        }
    }
}

template <int MODE, int SR>
cudaError_t launch_t(const CUtensorMap &tmap, const DetectParams &p, cudaStream_t stream) {
    auto kern = fdf_detect_kernel<MODE, SR>;
    const size_t smem = detect_smem_bytes(MODE, SR, p.words_per_row);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const unsigned long long items = (unsigned long long)p.n_frames * p.strips_per_frame;
    if (items == 0 || items > 0x7fffffffull) return cudaErrorInvalidValue;
    // persistent grid: as many CTAs as can be resident at once (each loops over tickets)
    int dev = 0, sms = 0, per_sm = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem)) != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorInvalidConfiguration;
    unsigned long long grid = (unsigned long long)sms * (unsigned)per_sm;
    if (grid > items) grid = items;
    kern<<<(unsigned)grid, kThreads, smem, stream>>>(tmap, p);
    return cudaGetLastError();
}

}  // namespace

size_t detect_smem_bytes(int mode, int sr, uint32_t words_per_row) {
    const size_t tile = (size_t)tile_rows(sr) * kTileW;
    const size_t plane = mode == NMS_OFF ? 0 : (size_t)sr * kPlaneW * 2;
    const size_t queue = (size_t)sr * kTileW * 2;
    return 2 * tile + plane + queue + 128 + 2 * (size_t)out_rows(mode, sr) * words_per_row * 4;
}

cudaError_t launch_detect(int mode, int sr, const CUtensorMap &tmap, const DetectParams &p, cudaStream_t stream) {
#define FDF_CASE(M, S) \
    if (mode == M && sr == S) return launch_t<M, S>(tmap, p, stream);
    FDF_CASE(0, 16)
    FDF_CASE(0, 32)
    FDF_CASE(1, 16)
    FDF_CASE(1, 32)
    FDF_CASE(2, 16)
    FDF_CASE(2, 32)
#undef FDF_CASE
    return cudaErrorInvalidValue;
}

cudaError_t launch_synth(uint8_t *d_frames, uint32_t n_frames, uint32_t w, uint32_t h, uint32_t pitch,
                         unsigned long long frame_stride, unsigned long long seed, uint32_t first_frame, uint32_t kind,
                         uint32_t amp, cudaStream_t stream) {
    const unsigned long long total = (unsigned long long)n_frames * h * ((w + 3u) / 4u);
    if (total == 0) return cudaSuccess;
    unsigned long long blocks = (total + 255ull) / 256ull;
    if (blocks > 148ull * 64ull) blocks = 148ull * 64ull;
    fdf_synth_kernel<<<(unsigned)blocks, 256, 0, stream>>>(d_frames, n_frames, w, h, pitch, frame_stride, seed,
                                                           first_frame, kind, amp);
    return cudaGetLastError();
}

}  // namespace fdf
