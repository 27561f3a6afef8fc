// fdf_kernels.cu -- sm_100a kernels of the FAST-n detection path.
//
// One persistent kernel does the whole path for a batch of frames (replaces fast_simd.rs:301-620):
//
//   work item  = (frame, strip of full-width rows), handed out through an atomic ticket in row-major
//                order, which is what makes the decoupled look-back deadlock-free.
//   8 warps, per chunk of a strip (two block barriers per chunk):
//     TMA 3-D tiled load (u8 tile 256 x (SR+6), zero-filled outside the image, double buffered,
//     issued two chunks ahead -- across strip boundaries --, completion on an mbarrier)  -> shared memory
//     phase A  : dense SWAR filter, 16 centres per thread (LDS.128 + PRMT + VABSDIFF4 + LOP3),
//                survivors pushed to the CTA's candidate queue                 (fast_simd.rs:368-520)
//                (the NMS pass of the previous chunk runs in the same barrier interval)
//     phase B  : one thread per candidate: 16 ring bytes -> brighter/darker 16-bit masks ->
//                rotate-AND arc test -> score in registers -> tagged score plane + keypoint list
//                                                                     (fast_simd.rs:115-297, 623-749)
//     NMS pass : strict 3x3 maximum on the shared-memory score plane          (fast_simd.rs:588-616)
//                survivors set one bit in the strip's shared-memory bit plane
//   per finished strip: popcount of the bit plane (warp + block prefix sums), then the bits are expanded
//     to (x, y) points, in row-major order, into a bump-allocated run of the staging buffer;
//     (count, position) is recorded.  The next strip's first tiles are already in flight.
//
// Two small kernels finish the ordered compaction (fast_simd.rs:550, 596-613: output is row-major):
//   fdf_scan_kernel   : exclusive prefix sum of the per-strip counts in (frame, strip) order -- block scan
//                       + decoupled look-back between scan tiles -- giving every strip's final offset and
//                       the CSR frame offsets;
//   fdf_gather_kernel : one warp per strip copies its run from staging to that offset.
// (Doing the look-back inside the detection kernel was measured at +45 % kernel time: with ~450 strips
// in flight every strip ends up waiting for all in-flight predecessors.)
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
    static constexpr int plane_bytes = (MODE == NMS_OFF) ? 0 : SR * kTileW * 2;  // u16: tag << 12 | score
    static constexpr int queue_off = plane_off + plane_bytes;
    static constexpr int queue_bytes = kQueueCap * 2;
    static constexpr int klist_off = queue_off + queue_bytes;
    static constexpr int klist_bytes = (MODE == NMS_OFF) ? 0 : 2 * kKlistCap * 2;  // two lists (chunk parity)
    static constexpr int wq_off = klist_off + klist_bytes;             // per-warp stage-1 -> stage-2 queues
    static constexpr int wq_bytes = kComputeWarps * kWarpQueueCap * 2;
    static constexpr int vtab_off = wq_off + wq_bytes;                  // validity tables: first / middle / last chunk
    static constexpr int vtab_bytes = 3 * kVtabWords * 4;
    static constexpr int misc_off = vtab_off + vtab_bytes;
    static constexpr int misc_bytes = 128;
    static constexpr int bits_off = misc_off + misc_bytes;  // bit plane: out_rows x words_per_row words (+ pad to 4)
    static_assert(bits_off % 16 == 0, "the bit plane is walked with 128-bit loads");
    static_assert(tile_bytes % 128 == 0, "TMA destination must stay 128-byte aligned");
    static_assert(SR % 16 == 0 && SR <= 64, "phase A walks rows in steps of 16; queue entries hold 6 row bits");
    static_assert((SR / 16) * 32 <= kWarpQueueCap, "a warp queue must hold every group stage 1 looks at");
    static_assert(vtab_off % 16 == 0, "the validity tables are read with 128-bit loads");
    static_assert(kGroupRows * kTileW <= kQueueCap, "a row group must always fit the candidate queue");
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
__global__ void __launch_bounds__(kThreads, SR >= 64 ? 2 : 4)
fdf_detect_kernel(const __grid_constant__ CUtensorMap tmap, const DetectParams p) {
    using L = Layout<MODE, SR>;
    constexpr int HS = (MODE == NMS_OFF) ? 0 : 1;  // score halo (rows and columns) needed by the 3x3 NMS
    constexpr int OUT_R = out_rows(MODE, SR);

    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *tiles = smem;
    uint16_t *plane = reinterpret_cast<uint16_t *>(smem + L::plane_off);
    uint16_t *queue = reinterpret_cast<uint16_t *>(smem + L::queue_off);
    uint16_t *klists = reinterpret_cast<uint16_t *>(smem + L::klist_off);               // [2][kKlistCap]
    uint16_t *wq = reinterpret_cast<uint16_t *>(smem + L::wq_off) + (threadIdx.x >> 5) * kWarpQueueCap;
    uint32_t *vtabs = reinterpret_cast<uint32_t *>(smem + L::vtab_off);                 // [3][kVtabWords]
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + L::misc_off);             // [2] tile landed
    uint32_t *qcount = reinterpret_cast<uint32_t *>(smem + L::misc_off + 16);          // [2] queue fill (chunk parity)
    uint32_t *kcount = reinterpret_cast<uint32_t *>(smem + L::misc_off + 24);          // [2] keypoints of the chunk
    uint32_t *s_ticket = reinterpret_cast<uint32_t *>(smem + L::misc_off + 32);        // [2] next strip's ticket
    uint32_t *warp_sums = reinterpret_cast<uint32_t *>(smem + L::misc_off + 48);       // [8]
    unsigned long long *s_base = reinterpret_cast<unsigned long long *>(smem + L::misc_off + 80);
    uint32_t *bits = reinterpret_cast<uint32_t *>(smem + L::bits_off);                  // [OUT_R][words_per_row]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int W = (int)p.w, H = (int)p.h;
    const int NC = (int)p.chunks_per_strip;
    const int WW = (int)p.words_per_row;
    const int nwords = OUT_R * WW;
    const int nunits = (nwords + 3) / 4;  // the bit plane in 128-bit units (padding words stay zero)
    const uint32_t total_items = p.n_frames * p.strips_per_frame;

    if (tid == 0) {
        tma_prefetch_desc(&tmap);
        mbar_init(&full_bar[0], 1);
        mbar_init(&full_bar[1], 1);
        fence_mbar_init();
        qcount[0] = qcount[1] = 0u;
        kcount[0] = kcount[1] = 0u;
        s_ticket[0] = atomicAdd(p.ticket, 1u);
    }
    for (int i = tid; i < 4 * nunits; i += kThreads) bits[i] = 0u;
    if (tid < 3 * kVtabWords) vtabs[tid] = valid_word<MODE>(W, vtab_chunk(tid / kVtabWords, NC), tid % kVtabWords);
    __syncthreads();

    const int t = (int)p.threshold, n = (int)p.count;
    const uint32_t kbias = filter_kbias(p.threshold);
    const int ahead = NC >= 2 ? 2 : 1;  // tiles requested this many chunks ahead (never beyond the next strip)
    // `nxt` is the ticket of the strip after `cur`.  Thread 0 draws it when the tile look-ahead first needs
    // it and publishes it through s_ticket[1 - parity]; everybody picks it up at the end of the strip.
    uint32_t cur = s_ticket[0], nxt = 0xffffffffu;
    bool have_nxt = false;
    uint32_t gc = 0;  // chunks processed by this CTA so far: tile stage = gc & 1, mbarrier parity = (gc >> 1) & 1

    // request the tile of chunk c of the current strip (c < NC) or of chunk c - NC of the next strip
    auto request_tile = [&](int c, uint32_t stream_index) {
        uint32_t item = cur;
        if (c >= NC) {
            if (c - NC >= NC) return;
            if (!have_nxt) {
                nxt = atomicAdd(p.ticket, 1u);
                have_nxt = true;
            }
            item = nxt;
            c -= NC;
        }
        if (item >= total_items) return;
        const uint32_t frame = item / p.strips_per_frame;
        const uint32_t strip = item - frame * p.strips_per_frame;
        const int ty0 = first_out_row(MODE) + (int)strip * OUT_R - HS - 3;  // image row of tile row 0
        const uint32_t stage = stream_index & 1u;
        mbar_expect_tx(&full_bar[stage], (uint32_t)L::tile_bytes);
        tma_load_3d(tiles + stage * L::tile_bytes, &tmap, c * kChunkW - kTileLead, ty0, (int)frame, &full_bar[stage]);
    };
    if (tid == 0)
        for (int c = 0; c < ahead; c++) request_tile(c, (uint32_t)c);

    for (uint32_t it = 0; cur < total_items; it++) {
        const uint32_t frame = cur / p.strips_per_frame;
        const uint32_t strip = cur - frame * p.strips_per_frame;
        bool nms_pending = false;  // the NMS pass of the previous chunk still has to run
        uint32_t pend_kn = 0;

        for (int c = 0; c < NC; c++, gc++) {
            const uint32_t stage = gc & 1u, cp = gc & 1u;
            const uint8_t *tile = tiles + stage * L::tile_bytes;
            const ChunkGeo g = make_geo<MODE>(W, H, WW, (int)strip, c, SR);
            const uint32_t tag = (uint32_t)(c % kTagPeriod) + 1u;
            if (MODE != NMS_OFF && c == 0) {  // strip start: restart the tag sequence on a cleared plane
                uint4 *pz = reinterpret_cast<uint4 *>(plane) + tid;
#pragma unroll
                for (int i = 0; i < L::plane_bytes / 16 / kThreads; i++) pz[i * kThreads] = make_uint4(0u, 0u, 0u, 0u);
            }
            if (MODE != NMS_OFF && nms_pending) {  // overlaps with this chunk's phase A (other warps)
                const ChunkGeo gp = make_geo<MODE>(W, H, WW, (int)strip, c - 1, SR);
                nms_list<MODE, SR>(tid, pend_kn, klists + (cp ^ 1u) * kKlistCap, plane, bits, gp,
                                   (uint32_t)((c - 1) % kTagPeriod) + 1u);
                nms_pending = false;
            }
            mbar_wait(&full_bar[stage], (gc >> 1) & 1u, p.flags);
            const uint32_t *vtab = vtabs + vtab_variant(c, NC) * kVtabWords;
            phase_a_warp<MODE, SR>(warp, lane, tile, wq, vtab, queue, &qcount[cp], g, kbias, 0, SR);
            __syncthreads();  // B1: queue complete; previous chunk fully suppressed

            const uint32_t qn = qcount[cp];
            if (tid == 0) {
                qcount[cp ^ 1u] = 0u;
                kcount[cp ^ 1u] = 0u;
            }
            if (MODE != NMS_OFF && c != 0 && tag == 1u) {  // every 15 chunks: restart the tags
                uint4 *pz = reinterpret_cast<uint4 *>(plane) + tid;
#pragma unroll
                for (int i = 0; i < L::plane_bytes / 16 / kThreads; i++) pz[i * kThreads] = make_uint4(0u, 0u, 0u, 0u);
                __syncthreads();
            }
            bool dense = false;
            if (qn <= (uint32_t)kQueueCap) {
                phase_b<MODE, SR>(tid, qn, tile, queue, plane, klists + cp * kKlistCap, &kcount[cp], bits, g, t, n, tag);
            } else {  // very dense content: redo the chunk kGroupRows rows at a time
                dense = true;
                for (int lo = 0; lo < SR; lo += kGroupRows) {
                    __syncthreads();
                    if (tid == 0) qcount[cp] = 0u;
                    __syncthreads();
                    phase_a_warp<MODE, SR>(warp, lane, tile, wq, vtab, queue, &qcount[cp], g, kbias, lo, lo + kGroupRows);
                    __syncthreads();
                    phase_b<MODE, SR>(tid, qcount[cp], tile, queue, plane, klists + cp * kKlistCap, &kcount[cp], bits, g,
                                      t, n, tag);
                }
            }
            __syncthreads();  // B2: tile[stage] is free again; every score of this chunk is in the plane

            if (tid == 0) request_tile(c + ahead, gc + (uint32_t)ahead);
            if (MODE != NMS_OFF) {
                const uint32_t kn = kcount[cp];
                if (dense || kn > (uint32_t)kKlistCap) {
                    nms_dense<MODE, SR>(tid, plane, bits, g, tag);
                    __syncthreads();  // the plane may be re-tagged / the counters reused
                } else {
                    nms_pending = true;
                    pend_kn = kn;
                }
            }
        }
        if (MODE != NMS_OFF && nms_pending) {
            const ChunkGeo gp = make_geo<MODE>(W, H, WW, (int)strip, NC - 1, SR);
            nms_list<MODE, SR>(tid, pend_kn, klists + ((gc - 1u) & 1u) * kKlistCap, plane, bits, gp,
                               (uint32_t)((NC - 1) % kTagPeriod) + 1u);
        }
        if (tid == 0) {
            if (!have_nxt) nxt = atomicAdd(p.ticket, 1u);  // (only when the look-ahead never reached the next strip)
            s_ticket[(it + 1u) & 1u] = nxt;
            have_nxt = false;
        }
        __syncthreads();  // strip end: all bits of the strip are set; plane and lists are free

        // ---- strip -> ordered run of points in the staging buffer --------------------------------------
        // the bit plane is walked in 128-bit units: one unit (4 words = 128 columns) per lane and step
        const int y0 = first_out_row(MODE) + (int)strip * OUT_R;  // first row this strip emits
        const EmitRange er = emit_range(warp, nunits);
        uint4 *bits4 = reinterpret_cast<uint4 *>(bits);
        uint32_t cnt = 0;
        for (int u = er.begin + lane; u < er.end; u += 32) {
            const uint4 v = bits4[u];
            cnt += (uint32_t)(__popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w));
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, d);  // warp total
        if (lane == 0) warp_sums[warp] = cnt;
        __syncthreads();
        if (warp == 0) {
            const uint32_t ws = lane < kThreads / 32 ? warp_sums[lane] : 0u;
            uint32_t wincl = ws;
#pragma unroll
            for (int d = 1; d < kThreads / 32; d <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, wincl, d);
                if (lane >= d) wincl += v;
            }
            if (lane < kThreads / 32) warp_sums[lane] = wincl - ws;  // exclusive offset of each warp
            if (lane == kThreads / 32 - 1) {
                unsigned long long o = 0ull;
                if (wincl != 0u) o = atomicAdd(p.cursor, (unsigned long long)wincl);  // the strip's run in staging
                *s_base = o;
                p.item_count[cur] = wincl;
                p.item_src[cur] = o;
            }
        }
        __syncthreads();
        if (cnt != 0u) {  // warp-uniform: this warp's range holds keypoints
            unsigned long long o = *s_base + warp_sums[warp];
            for (int base = er.begin; base < er.end; base += 32) {
                const int u = base + lane;
                uint4 v = make_uint4(0u, 0u, 0u, 0u);
                if (u < er.end) v = bits4[u];
                const uint32_t c = (uint32_t)(__popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w));
                uint32_t incl = c;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t w = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= d) incl += w;
                }
                if (c != 0u) {
                    unsigned long long oo = o + (incl - c);
                    const int wi = 4 * u;
                    int row = wi / WW, col = wi - row * WW;
                    const uint32_t words[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        if (words[k] != 0u) {
                            emit_word(words[k], (uint32_t)col * 32u, (uint32_t)(y0 + row), oo, p.cap, p.staging);
                            oo += (unsigned long long)__popc(words[k]);
                        }
                        if (++col == WW) {
                            col = 0;
                            row++;
                        }
                    }
                    bits4[u] = make_uint4(0u, 0u, 0u, 0u);  // leave the plane zeroed for the next strip
                }
                o += __shfl_sync(0xffffffffu, incl, 31);
            }
        }
        cur = s_ticket[(it + 1u) & 1u];
        // (the next strip's first bit is set after its first barrier, i.e. after every warp left this loop)
    }
}

// ---- ordered compaction, step 2: exclusive scan of the per-strip counts -------------------------------
// One scan tile = kScanTile consecutive strips; tiles take tickets in order and chain through the
// decoupled look-back above.  Also writes the CSR frame offsets.
__global__ void __launch_bounds__(kScanThreads) fdf_scan_kernel(const DetectParams p, uint32_t n_items) {
    __shared__ uint32_t s_tile;
    __shared__ uint32_t warp_sums[kScanThreads / 32];
    __shared__ unsigned long long s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(p.scan_ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t first = tile * kScanTile + (uint32_t)tid * kScanItemsPerThread;
    uint32_t v[kScanItemsPerThread];
    uint32_t sum = 0;
#pragma unroll
    for (int k = 0; k < kScanItemsPerThread; k++) {
        v[k] = first + k < n_items ? p.item_count[first + k] : 0u;
        sum += v[k];
    }
    uint32_t incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += u;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const uint32_t ws = lane < kScanThreads / 32 ? warp_sums[lane] : 0u;
        uint32_t wincl = ws;
#pragma unroll
        for (int d = 1; d < kScanThreads / 32; d <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, wincl, d);
            if (lane >= d) wincl += u;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, wincl, kScanThreads / 32 - 1);
        if (lane < kScanThreads / 32) warp_sums[lane] = wincl - ws;  // exclusive offset of each warp
        const unsigned long long excl = lookback(p.scan_status, tile, total, lane, p.flags);
        if (lane == 0) s_base = excl;
    }
    __syncthreads();
    unsigned long long o = s_base + warp_sums[warp] + (incl - sum);
#pragma unroll
    for (int k = 0; k < kScanItemsPerThread; k++) {
        const uint32_t i = first + k;
        if (i < n_items) {
            p.item_dst[i] = o;
            if (i % p.strips_per_frame == 0) p.offsets[i / p.strips_per_frame] = o;  // first strip of a frame
            if (i == n_items - 1) p.offsets[p.n_frames] = o + v[k];
        }
        o += v[k];
    }
}

// ---- ordered compaction, step 3: one warp per strip moves its run to its final position ------------------
__global__ void __launch_bounds__(256) fdf_gather_kernel(const DetectParams p, uint32_t n_items) {
    const uint32_t item = blockIdx.x * 8u + (threadIdx.x >> 5);
    const uint32_t lane = threadIdx.x & 31u;
    if (item >= n_items) return;
    const uint32_t cnt = p.item_count[item];
    const unsigned long long src = p.item_src[item], dst = p.item_dst[item];
    for (uint32_t k = lane; k < cnt; k += 32u)
        if (src + k < p.cap && dst + k < p.cap) p.out[dst + k] = p.staging[src + k];
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
    const size_t plane = mode == NMS_OFF ? 0 : (size_t)sr * kTileW * 2;
    const size_t queue = (size_t)kQueueCap * 2;
    const size_t klist = mode == NMS_OFF ? 0 : (size_t)2 * kKlistCap * 2;
    const size_t bit_words = (((size_t)out_rows(mode, sr) * words_per_row + 3) / 4) * 4;
    const size_t wq = (size_t)kComputeWarps * kWarpQueueCap * 2, vtab = (size_t)3 * (kTileW / 4) * 4;
    return 2 * tile + plane + queue + klist + wq + vtab + 128 + bit_words * 4;
}

cudaError_t launch_detect(int mode, int sr, const CUtensorMap &tmap, const DetectParams &p, cudaStream_t stream) {
#define FDF_CASE(M, S) \
    if (mode == M && sr == S) return launch_t<M, S>(tmap, p, stream);
    FDF_CASE(0, 32)
    FDF_CASE(0, 64)
    FDF_CASE(1, 32)
    FDF_CASE(1, 64)
    FDF_CASE(2, 32)
    FDF_CASE(2, 64)
#undef FDF_CASE
    return cudaErrorInvalidValue;
}

cudaError_t launch_scan(const DetectParams &p, cudaStream_t stream) {
    const unsigned long long items = (unsigned long long)p.n_frames * p.strips_per_frame;
    if (items == 0 || items > 0x7fffffffull) return cudaErrorInvalidValue;
    const unsigned tiles = (unsigned)((items + kScanTile - 1) / kScanTile);
    fdf_scan_kernel<<<tiles, kScanThreads, 0, stream>>>(p, (uint32_t)items);
    return cudaGetLastError();
}

cudaError_t launch_gather(const DetectParams &p, cudaStream_t stream) {
    const unsigned long long items = (unsigned long long)p.n_frames * p.strips_per_frame;
    if (items == 0 || items > 0x7fffffffull) return cudaErrorInvalidValue;
    fdf_gather_kernel<<<(unsigned)((items + 7) / 8), 256, 0, stream>>>(p, (uint32_t)items);
    return cudaGetLastError();
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
