// fdf_kernels.cu -- sm_100a kernels of the FAST-n detection path.
//
// One persistent kernel does the detection for a batch of frames (replaces fast_simd.rs:301-620):
//
//   work item  = (frame, strip of full-width rows), handed out through an atomic ticket; a strip is walked left to
//                right in chunks of 240 columns, each staged as a 256-byte-wide tile.
//   10 warps per CTA in two groups that are coupled ONLY through mbarriers (no CTA-wide barrier in the chunk loop):
//     4 filter warps  wait for the chunk's tile (TMA 3-D tiled load, u8 tile 256 x (SR+6), zero-filled outside the
//                     image, kTileStages buffers, completion on an mbarrier)
//                     phase A stage 1: every pixel, 16 per lane (LDS.128 + VABSDIFF4 + LOP3), north/south pair; groups
//                     with a survivor go to the warp's segment of the chunk's entry table (ballot, no atomics);
//                     one barrier among the filter warps;
//                     phase A stage 2: one lane per entry, the entries of ALL warps dealt evenly to the filter threads
//                     (a horizontal edge puts most of a chunk's entries into one warp's rows): adds the east/west pair
//                     (PRMT byte shifts) and pushes the surviving centres to the chunk's candidate queue
//                                                                                         (fast_simd.rs:368-520)
//     6 test warps    wait for the candidate queue
//                     phase B: one thread per candidate: 16 ring bytes -> one "dual" word per ring pixel -> best
//                     window (VIMNMX3.U16x2) = segment test AND MaxThreshold score -> tagged score plane + keypoint
//                     list                                                       (fast_simd.rs:115-297, 623-749)
//                     barrier among the test warps; one thread requests the tile of chunk k + kTileStages;
//                     NMS pass: strict 3x3 maximum on the raw plane cells per listed keypoint, survivors written
//                     directly into the chunk's run of the staging buffer            (fast_simd.rs:588-616)
//                     barrier; one thread closes the run (run record, strip total)
//   Every chunk leaves one unordered run of (row, x) entries in the staging buffer plus a run record.
//
// Two small kernels finish the ordered compaction (fast_simd.rs:550, 596-613: output is row-major):
//   fdf_scan_kernel   : exclusive prefix sum of the per-strip counts in (frame, strip) order -- block scan
//                       + decoupled look-back between scan tiles -- giving every strip's final offset and
//                       the CSR frame offsets;
//   fdf_gather_kernel : persistent CTAs, one strip at a time: scatters the strip's runs into a two-level bit plane in
//                       shared memory and expands it, row-major, to (x, y) points at the strip's final offset.  With
//                       at most kGatherThreads strips in the batch (one image) it also does the scan itself: no
//                       fdf_scan_kernel launch.
//   fdf_shard_push_kernel (sharded batches, one process per GPU): global CSR offsets from the all-gathered local ones,
//                       and the rank's points copied to their place in the batch result in rank 0's memory (NVLink).
// (Doing the look-back inside the detection kernel was measured at +45 % kernel time: with ~450 strips
// in flight every strip ends up waiting for all in-flight predecessors.  Keeping the strip bit plane inside
// the detection kernel cost 30 KB of shared memory per CTA, i.e. one resident CTA per SM.  Structural variants that
// were built, measured and rejected are kept as patches under tools/experiments/.)
#include "fdf_kernels.cuh"

#include <cstdlib>

#include "fdf_core.cuh"
#include "fdf_strip.cuh"
#include "fdf_synth.cuh"

namespace fdf {
namespace {

constexpr uint32_t kFlagLookbackTimeout = 1u;
constexpr uint32_t kFlagTmaTimeout = 2u;
constexpr uint32_t kFlagStagingOverflow = 4u;
constexpr unsigned long long kStatusAggregate = 1ull << 62;
constexpr unsigned long long kStatusPrefix = 2ull << 62;
constexpr unsigned long long kStatusValueMask = (1ull << 62) - 1ull;
constexpr uint32_t kSpinLimit = 1u << 22;
constexpr uint32_t kWaitHintNs = 100000u;  // mbarrier.try_wait suspend-time hint
constexpr unsigned long long kWaitLimitNs = 4000000000ull;  // 4 s

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


__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// try_wait suspends the thread (no issue slots used) until the phase completes or the time hint (ns) runs out
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(kWaitHintNs)
        : "memory");
    return ok != 0;
}

// Waits for the phase; a wait that lasts longer than kWaitLimitNs can only be a bug in the pipeline: it is turned
// into an error flag (the host reports FDF_ERR_INTERNAL) instead of a hung GPU.
// (`abort` is a flag in shared memory: once one wait of the CTA has timed out, no other wait of the CTA blocks, so
// that the kernel still ends quickly.)  A three-instruction PTX spin (probe, branch, count) and named hardware
// barriers instead of polled mbarriers were both measured: 1 % slower and +-0 (the polls use issue slots that
// nobody else wants), so the plain loop stays.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity, uint32_t *flags, volatile uint32_t *abort) {
    if (mbar_try_wait(bar, parity)) return;
    unsigned long long t0 = 0ull;
    for (uint32_t spins = 1;; spins++) {
        if (mbar_try_wait(bar, parity)) return;
        if ((spins & 63u) == 0u) {  // (rarely reached: a wait normally ends within a few rounds)
            if (*abort != 0u) return;
            const unsigned long long now = global_timer_ns();
            if (t0 == 0ull) t0 = now;
            if (now - t0 > kWaitLimitNs) {
                atomicOr(flags, kFlagTmaTimeout);
                *abort = 1u;
                return;
            }
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

// TMA: prefetch a tile into L2 only (no shared memory, no completion): the later load of the same box then hits L2
__device__ __forceinline__ void tma_prefetch_l2_3d(const CUtensorMap *tmap, int x, int y, int z) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(
                     reinterpret_cast<uint64_t>(tmap)),
                 "r"(x), "r"(y), "r"(z)
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
#ifndef FDF_TILE_STAGES
#define FDF_TILE_STAGES 2
#endif
#ifndef FDF_L2_PREFETCH
#define FDF_L2_PREFETCH 1
#endif
constexpr int kTileStages = FDF_TILE_STAGES;  // tile buffers per CTA: the tile of chunk k + kTileStages is requested when
                                              // phase B of chunk k is done, so TMA latency (~1700 cycles under load)
                                              // is hidden behind kTileStages - 1 chunks of work
constexpr int kQueueBufs = kTileStages;       // candidate queues: the filter may run kTileStages - 1 chunks ahead of the test
static_assert(kTileStages >= 2 && kTileStages <= 4, "misc layout holds up to 4 barriers of each kind");

struct LayoutSizes {
    int tile_bytes, plane_off, plane_bytes, queue_off, klist_off, ent_off, vtab_off, misc_off, total;
};
__host__ __device__ constexpr LayoutSizes layout_sizes(int sr) {
    LayoutSizes l = {};
    l.tile_bytes = tile_rows(sr) * kTileW;                       // one TMA box
    l.plane_off = kTileStages * l.tile_bytes;
    l.plane_bytes = sr * kPlaneW * 2;                            // u16: tag << 12 | score (Off mode: score 1)
    l.queue_off = l.plane_off + l.plane_bytes;                   // candidate queues [kQueueBufs][kQueueCap] u16
    l.klist_off = l.queue_off + kQueueBufs * kQueueCap * 2;      // keypoint list of the chunk being tested [kQueueCap] u16
    l.ent_off = l.klist_off + kQueueCap * 2;                     // stage-1 entry tables [2][warps][kWarpQueueCap] u8
    l.vtab_off = l.ent_off + 2 * kFilterWarps * kWarpQueueCap;   // validity tables: first / middle / last chunk
    l.misc_off = l.vtab_off + 3 * kVtabWords * 4;
    l.total = l.misc_off + 192;
    return l;
}

// misc block (byte offsets)
constexpr int kMiscFullBar = 0;     // [4] u64 tile landed
constexpr int kMiscQFull = 32;      // [4] u64 candidate queue complete (one arrival per filter warp)
constexpr int kMiscQCount = 64;     // [4] u32 candidate queue fill
constexpr int kMiscTicket = 80;     // [2] u32 strip tickets (strip parity)
constexpr int kMiscSCount = 88;     // u32 keypoints staged by the chunk so far
constexpr int kMiscSTotal = 92;     // u32 keypoints of the strip so far
constexpr int kMiscSBase = 96;      // u64 where the chunk's run starts in the staging buffer
constexpr int kMiscSBlock = 104;    // [2] u64 the CTA's staging block: next free entry, end
constexpr int kMiscAbort = 120;     // u32 a wait timed out
constexpr int kMiscKCount = 124;    // [2] u32 keypoint list fill (chunk parity)
constexpr int kMiscNEnt = 132;      // [2][4] u32 stage-1 entries per filter warp (chunk parity)
constexpr int kMiscDropped = 164;   // u32 a staged entry did not fit the staging buffer
constexpr int kMiscTrace = 168;     // i32 (trace builds) index of this CTA in the trace table

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

// Timeline trace (tools/trace_timeline.py builds the library with -DFDF_TRACE): the CTAs resident on SM 0 write
// clock64() at every phase boundary of their first kTraceChunks chunks, per warp (lane 0), into a global table.
#ifdef FDF_TRACE
constexpr int kTraceCtas = 4, kTraceChunks = 200, kTraceSlots = 12;
__device__ long long g_trace[kTraceCtas][16][kTraceChunks][kTraceSlots];
__device__ unsigned int g_trace_n;
#define FDF_CLK(slot)                                                                                           \
    if (trace_cta >= 0 && lane == 0 && gc < (uint32_t)kTraceChunks) g_trace[trace_cta][warp][gc][slot] = clock64();
#else
#define FDF_CLK(slot)
#endif

// ---- the detection kernel ----------------------------------------------------------------------
// Two groups of warps per CTA, coupled only through mbarriers (no CTA-wide barrier inside the chunk loop):
//   filter warps (0 .. kFilterWarps-1): wait for chunk k's tile, run phase A into candidate queue k % kQueueBufs
//       (stage 1 per warp -> barrier among the filter warps -> stage 2 over all warps' entries), arrive on q_full,
//       go on to chunk k + 1;
//   test warps (the others): wait on q_full, run phase B, barrier among themselves, then one thread requests the tile
//       of chunk k + kTileStages (tile buffer and candidate queue of chunk k are free now), all run the NMS pass over
//       the chunk's keypoint list, barrier, and one thread closes the chunk's run while the others already wait for
//       chunk k + 1.
// Which buffer is protected by what:
//   tile[s], queue[s] (s = k % kTileStages): written by TMA / the filter warps after full_bar[s] of chunk k; read by
//       the test warps until their barrier after phase B of chunk k; the tile of chunk k + kTileStages is requested
//       after that barrier.  A landed tile therefore implies that its queue is free again.
//   entry table [k & 1]: written by stage 1 of chunk k before the filter barrier, read by stage 2 after it; chunk
//       k + 2 writes it again after the filter barrier of chunk k + 1, which every filter warp reaches only after its
//       stage 2 of chunk k.
//   plane: cells carry a 4-bit chunk tag, larger tags are newer; written in phase B, read in the NMS pass, both inside
//       one test-group barrier interval each; cleared every kTagPeriod chunks.
//   klist, scount, s_base: written / read by the test warps between their two barriers; reset by the thread that
//       closes the run, before the barrier that opens the next chunk's phase B ... (kcount is double-buffered by chunk
//       parity because phase B of chunk k + 1 appends while that thread still reads chunk k's count).
__device__ __forceinline__ void bar_test_group() {
    asm volatile("bar.sync 1, %0;" ::"n"(kTestThreads) : "memory");
}
__device__ __forceinline__ void bar_filter_group() {
    asm volatile("bar.sync 2, %0;" ::"n"(kFilterThreads) : "memory");
}

// Staging space is handed out in two levels: a CTA takes blocks of kStageBlock entries from the global cursor
// (one contended atomic every few dozen chunks instead of one per chunk, which sat on the test warps' critical
// path) and cuts its runs from the current block.  s_block[0] = next free entry, s_block[1] = end of the block.
// Only thread t0 calls these, between barriers of the test group.
// A chunk's run is opened before its size is known (room for a full keypoint list), written by the NMS pass, and
// closed at its real size: the unused tail goes back to the block.
__device__ __forceinline__ unsigned long long reserve_staging(uint32_t count, unsigned long long *s_block,
                                                              const DetectParams &p) {
    unsigned long long next = s_block[0];
    if (next + count > s_block[1]) {
        const unsigned long long n = count > (uint32_t)kStageBlock ? count : (unsigned long long)kStageBlock;
        next = atomicAdd(p.cursor, n);
        s_block[1] = next + n;
    }
    s_block[0] = next + count;
    return next;
}

__device__ __forceinline__ unsigned long long open_run(unsigned long long *s_block, const DetectParams &p) {
    if (s_block[0] + (unsigned long long)kQueueCap > s_block[1]) {
        s_block[0] = atomicAdd(p.cursor, (unsigned long long)kStageBlock);
        s_block[1] = s_block[0] + (unsigned long long)kStageBlock;
    }
    return s_block[0];
}

template <int MODE, int SR>
__global__ void __launch_bounds__(kThreads, SR >= 48 ? 3 : 4)
fdf_detect_kernel(const __grid_constant__ CUtensorMap tmap, const DetectParams p) {
    if (p.idle_sm_stride != 0u) {  // SMs kept free for the exchange kernels of a sharded batch (fdf_set_idle_sms)
        unsigned int smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        if (smid % p.idle_sm_stride == p.idle_sm_stride - 1u) return;  // (block-uniform; the strips go by ticket)
    }
    // A small grid (one image) leaves most of the GPU empty: let the grid behind it (the gather kernel, launched with
    // programmatic stream serialisation) be placed right away; it waits at its griddepcontrol.wait for this grid's end.
    if (SR == 32 && gridDim.x < 2u * 148u) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    constexpr LayoutSizes L = layout_sizes(SR);
    constexpr int HS = (MODE == NMS_OFF) ? 0 : 1;  // score halo (rows and columns) needed by the 3x3 NMS
    constexpr int OUT_R = out_rows(MODE, SR);
    static_assert(L.tile_bytes % 128 == 0, "TMA destination must stay 128-byte aligned");
    static_assert(L.plane_bytes % 16 == 0 && L.plane_off % 16 == 0, "the plane is cleared with 128-bit stores");
    static_assert(SR % (2 * kFilterWarps) == 0 && SR <= 64, "filter warps take row pairs; queue entries hold 6 row bits");
    static_assert(L.vtab_off % 16 == 0 && L.misc_off % 8 == 0, "alignment of the validity tables / mbarriers");

    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *tiles = smem;
    uint16_t *plane = reinterpret_cast<uint16_t *>(smem + L.plane_off);
    uint16_t *queues = reinterpret_cast<uint16_t *>(smem + L.queue_off);              // [kQueueBufs][kQueueCap]
    uint16_t *klist = reinterpret_cast<uint16_t *>(smem + L.klist_off);               // [kQueueCap]
    uint8_t *ents = smem + L.ent_off;                                                 // [2][kFilterWarps][kWarpQueueCap]
    uint32_t *vtabs = reinterpret_cast<uint32_t *>(smem + L.vtab_off);                 // [3][kVtabWords]
    uint8_t *misc = smem + L.misc_off;
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(misc + kMiscFullBar);
    uint64_t *q_full = reinterpret_cast<uint64_t *>(misc + kMiscQFull);
    uint32_t *qcount = reinterpret_cast<uint32_t *>(misc + kMiscQCount);
    uint32_t *s_ticket = reinterpret_cast<uint32_t *>(misc + kMiscTicket);
    uint32_t *scount = reinterpret_cast<uint32_t *>(misc + kMiscSCount);
    uint32_t *s_total = reinterpret_cast<uint32_t *>(misc + kMiscSTotal);
    unsigned long long *s_base = reinterpret_cast<unsigned long long *>(misc + kMiscSBase);
    unsigned long long *s_block = reinterpret_cast<unsigned long long *>(misc + kMiscSBlock);
    volatile uint32_t *s_abort = reinterpret_cast<volatile uint32_t *>(misc + kMiscAbort);
    uint32_t *kcount = reinterpret_cast<uint32_t *>(misc + kMiscKCount);
    uint32_t *s_nent = reinterpret_cast<uint32_t *>(misc + kMiscNEnt);
    uint32_t *s_dropped = reinterpret_cast<uint32_t *>(misc + kMiscDropped);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int W = (int)p.w, H = (int)p.h;
    const int NC = (int)p.chunks_per_strip;
    // A work item is one strip, or -- for small batches, where whole strips leave most CTAs idle or make a long tail --
    // one of p.parts equal column ranges of a strip (NCI chunks each; the host only splits evenly and keeps
    // NCI >= kTileStages).  Chunks are independent, so nothing but the ticket arithmetic changes.
    // Only the 32-row kernels (the ones small inputs run) carry the split: for the others `parts` is the constant 1 and
    // the arithmetic below folds away -- it cost the large-batch kernel 2.5 % when it was there at run time.
    const uint32_t parts = SR == 32 ? p.parts : 1u;
    const int NCI = NC / (int)parts;
    const uint32_t total_items = p.n_frames * p.strips_per_frame * parts;
    const bool is_filter = warp < kFilterWarps;
    const int ttid = tid - kFilterThreads;  // thread index inside the test group
    // the thread that draws tickets, requests tiles and keeps the run records: lane 0 of the LAST test warp, which gets
    // the smallest share of every candidate / keypoint list -- its serial work between the group's barriers then
    // overlaps the other warps' list work instead of extending the critical path
    const bool t0 = ttid == 32 * (kTestWarps - 1);

    uint32_t cur = 0u, nxt = 0xffffffffu;
    bool have_nxt = false;
    const int ahead = NCI >= kTileStages ? kTileStages : NCI;  // tiles requested this many chunks ahead (never beyond the next item)

    // request the tile of chunk c of the current item (c < NCI) or of chunk c - NCI of the next item; `it` is the
    // current item's sequence number in this CTA.  When the work is exhausted the barrier is completed without
    // a tile, so that the filter warps wake up and see the end ticket.
    auto request_tile = [&](int c, uint32_t stream_index, uint32_t it) {
        uint32_t item = cur;
        if (c >= NCI) {
            if (c - NCI >= NCI) return;
            if (!have_nxt) {
                nxt = atomicAdd(p.ticket, 1u);
                have_nxt = true;
                s_ticket[(it + 1u) & 1u] = nxt;
            }
            item = nxt;
            c -= NCI;
        }
        const uint32_t stage = stream_index % (uint32_t)kTileStages;
        if (item >= total_items) {
            mbar_arrive(&full_bar[stage]);
            return;
        }
        const uint32_t sg = item / parts;  // strip in (frame, strip) order
        c += (int)(item - sg * parts) * NCI;  // chunk of the strip
        const uint32_t frame = sg / p.strips_per_frame;
        const uint32_t strip = sg - frame * p.strips_per_frame;
        const int ty0 = first_out_row(MODE) + (int)strip * OUT_R - HS - 3;  // image row of tile row 0
        mbar_expect_tx(&full_bar[stage], (uint32_t)L.tile_bytes);
        tma_load_3d(tiles + stage * L.tile_bytes, &tmap, c * kChunkW - kTileLead, ty0, (int)frame, &full_bar[stage]);
        // and pull the strip's next tile into L2 (cp.async.bulk.prefetch.tensor: no shared memory, no completion), so
        // that its load, one chunk from now, is an L2 hit (+1 %; 2 or 4 chunks ahead measured no better)
        if (FDF_L2_PREFETCH > 0 && c + FDF_L2_PREFETCH < NC)
            tma_prefetch_l2_3d(&tmap, (c + FDF_L2_PREFETCH) * kChunkW - kTileLead, ty0, (int)frame);
    };

    if (t0) {
        tma_prefetch_desc(&tmap);
        for (int i = 0; i < kTileStages; i++) mbar_init(&full_bar[i], 1);
        for (int i = 0; i < kQueueBufs; i++) {
            mbar_init(&q_full[i], kFilterWarps);
            qcount[i] = 0u;
        }
        fence_mbar_init();
        *scount = 0u;
        *s_total = 0u;
        *s_abort = 0u;
        *s_dropped = 0u;
        kcount[0] = kcount[1] = 0u;
        s_block[0] = s_block[1] = 0ull;
        *s_base = open_run(s_block, p);
        cur = atomicAdd(p.ticket, 1u);
        s_ticket[0] = cur;
    }
    if (tid < 3 * kVtabWords) vtabs[tid] = valid_group_mask<MODE>(W, vtab_chunk(tid / kVtabWords, NC), tid % kVtabWords);
    auto clear_plane = [&](int i0, int n) {  // by n threads, i0 = index of this one
        uint4 *pz = reinterpret_cast<uint4 *>(plane);
        for (int i = i0; i < L.plane_bytes / 16; i += n) pz[i] = make_uint4(0u, 0u, 0u, 0u);
    };
    clear_plane(tid, kThreads);
#ifdef FDF_TRACE
    volatile int &s_trace_cta = *reinterpret_cast<volatile int *>(misc + kMiscTrace);
    if (tid == 0) {
        unsigned int smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        s_trace_cta = -1;
        if (smid == 0u) {
            const unsigned int k = atomicAdd(&g_trace_n, 1u);
            if (k < (unsigned)kTraceCtas) s_trace_cta = (int)k;
        }
    }
#endif
    __syncthreads();
#ifdef FDF_TRACE
    const int trace_cta = s_trace_cta;
#endif
    cur = s_ticket[0];
    if (t0)
        for (int c = 0; c < ahead; c++) request_tile(c, (uint32_t)c, 0u);  // (ahead <= NCI: all of the first item)

    const int t = (int)p.threshold, n = (int)p.count;
    uint32_t gc = 0;   // chunks processed by this CTA so far
    uint32_t qb = 0, qpar = 0;  // tile stage = queue buffer = gc % kTileStages, and the parity of their mbarrier phases

    if (is_filter) {
        // ================================ filter warps =================================================
        const uint32_t kbias = filter_kbias(p.threshold);
        for (uint32_t it = 0; cur < total_items; it++) {
            const uint32_t sg = cur / parts;
            const int c0 = (int)(cur - sg * parts) * NCI;
            const uint32_t frame = sg / p.strips_per_frame;
            const uint32_t strip = sg - frame * p.strips_per_frame;
            for (int ci = 0; ci < NCI; ci++, gc++) {
                const int c = c0 + ci;  // chunk of the strip
                const ChunkGeo g = make_geo<MODE>(W, H, (int)strip, c, SR);
                const uint8_t *tile = tiles + qb * L.tile_bytes;
                mbar_wait(&full_bar[qb], qpar, p.flags, s_abort);  // the tile has landed (also: queue qb is free again)
                FDF_CLK(0)
                uint8_t *eb = ents + (gc & 1u) * (kFilterWarps * kWarpQueueCap);
                uint32_t *nent = s_nent + (gc & 1u) * 4u;
                const uint32_t ne = phase_a_stage1<MODE, SR, kFilterWarps>(warp, lane, tile, eb + warp * kWarpQueueCap, g, kbias);
                if (lane == 0) nent[warp] = ne;
                FDF_CLK(9)
                bar_filter_group();  // every warp's entries of this chunk are in the table
                phase_a_stage2<MODE, SR, kFilterWarps>(tid, kFilterThreads, tile, eb, nent,
                                                       vtabs + vtab_variant(c, NC) * kVtabWords, queues + qb * kQueueCap,
                                                       &qcount[qb], kbias);
                __syncwarp();
                if (lane == 0) mbar_arrive(&q_full[qb]);
                FDF_CLK(1)
                if (++qb == (uint32_t)kQueueBufs) {
                    qb = 0;
                    qpar ^= 1u;
                }
            }
            // the next strip's ticket is published before its first tile is requested (or the end is signalled)
            mbar_wait(&full_bar[qb], qpar, p.flags, s_abort);
            cur = s_ticket[(it + 1u) & 1u];
        }
        return;
    }

    // ==================================== test warps ===================================================
    uint32_t tag = 1u;  // (gc % kTagPeriod) + 1
    const int twarp = warp - kFilterWarps;
    // t0, between the group's barriers: the chunk's run record and the strip's running total
    auto close_run = [&](uint32_t slot, uint32_t count, bool last) {
        if (count != 0u) p.run_base[slot] = *s_base;
        p.run_count[slot] = count;
        const uint32_t tot = *s_total + count;
        if (last) {  // (a strip cut into parts: the counts of its items add up; the host zeroed them)
            if (parts == 1u) p.item_count[cur] = tot;
            else if (tot != 0u) atomicAdd(&p.item_count[cur / parts], tot);
        }
        *s_total = last ? 0u : tot;
    };
    for (uint32_t it = 0; cur < total_items; it++) {
        const uint32_t sg = cur / parts;
        const int c0 = (int)(cur - sg * parts) * NCI;
        const uint32_t frame = sg / p.strips_per_frame;
        const uint32_t strip = sg - frame * p.strips_per_frame;
        for (int ci = 0; ci < NCI; ci++, gc++) {
            const int c = c0 + ci;  // chunk of the strip
            const uint8_t *tile = tiles + qb * L.tile_bytes;
            uint16_t *queue = queues + qb * kQueueCap;
            const ChunkGeo g = make_geo<MODE>(W, H, (int)strip, c, SR);
            const uint32_t slot = sg * (uint32_t)NC + (uint32_t)c;
            if (tag == 1u && gc != 0u) {  // every 15 chunks: restart the tags on a cleared plane
                clear_plane(ttid, kTestThreads);
                bar_test_group();
            }
            mbar_wait(&q_full[qb], qpar, p.flags, s_abort);
            mbar_wait(&full_bar[qb], qpar, p.flags, s_abort);  // (completed long ago: makes the tile visible here too)
            FDF_CLK(4)
            const uint32_t qn = qcount[qb];
            bool dropped = false;
            if (qn <= (uint32_t)kQueueCap) {
                phase_b<MODE, SR>(ttid, lane, kTestThreads, qn, tile, queue, klist, &kcount[gc & 1u], plane, t, n, tag);
                FDF_CLK(5)
                bar_test_group();  // every score of this chunk is in the plane and its keypoint list is complete
                if (t0) {
                    qcount[qb] = 0u;
                    request_tile(ci + ahead, gc + (uint32_t)ahead, it);
                }
                FDF_CLK(6)
                // In the NMS modes the last test warp stays out of this pass: its lane 0 has just spent ~1000 cycles on the
                // tile request, and a chunk's ~170 keypoints are one block of 32 per warp -- with a share of its own it
                // would start that block late and hold up the barrier below (profiles/r02_v19_timeline_3ctas.txt, "last
                // test warp alone"; -0.6 %).  Off mode lists five times as many keypoints and needs all six (+3 % without).
                if (MODE == NMS_OFF)
                    dropped = emit_list<MODE, SR>(ttid, kTestThreads, kcount[gc & 1u], klist, plane, scount, *s_base,
                                                  p.staging_cap, p.staging, g);
                else if (twarp < kTestWarps - 1)
                    dropped = emit_list<MODE, SR>(ttid, kTestThreads - 32, kcount[gc & 1u], klist, plane, scount, *s_base,
                                                  p.staging_cap, p.staging, g);
                FDF_CLK(7)
                bar_test_group();  // the run is complete; the keypoint list is free
                if (t0) {
                    const uint32_t count = *scount;
                    *scount = 0u;
                    kcount[gc & 1u] = 0u;  // (next used two chunks from now)
                    close_run(slot, count, ci == NCI - 1);
                    s_block[0] = *s_base + count;  // give the unused tail back
                    *s_base = open_run(s_block, p);
                }
                FDF_CLK(8)
            } else {
                // very dense content (e.g. noise): the queue overflowed.  Every scored pixel of the chunk gets the full
                // test straight from the tile, then the score plane is scanned: count, reserve the run at its exact size,
                // write
                phase_b_dense<MODE, SR>(twarp, lane, kTestWarps, tile, plane, g, c, t, n, tag);
                bar_test_group();  // every score of this chunk is in the plane; tile and queue are free
                if (t0) {
                    qcount[qb] = 0u;
                    request_tile(ci + ahead, gc + (uint32_t)ahead, it);
                }
                nms_dense<MODE, SR>(ttid, kTestThreads, 0, plane, scount, 0ull, p.staging_cap, p.staging, g, tag);
                bar_test_group();
                const uint32_t kn = *scount;
                bar_test_group();
                if (t0) {
                    *scount = 0u;
                    s_block[0] = *s_base;  // the run opened for this chunk is not used: the exact size is known now
                    *s_base = reserve_staging(kn, s_block, p);
                }
                bar_test_group();
                if (kn != 0u)
                    dropped = nms_dense<MODE, SR>(ttid, kTestThreads, 1, plane, scount, *s_base, p.staging_cap, p.staging, g, tag);
                bar_test_group();
                if (t0) {
                    *scount = 0u;
                    close_run(slot, kn, ci == NCI - 1);
                    *s_base = open_run(s_block, p);
                }
            }
            if (dropped) *s_dropped = 1u;
            tag = tag == (uint32_t)kTagPeriod ? 1u : tag + 1u;
            if (++qb == (uint32_t)kQueueBufs) {
                qb = 0;
                qpar ^= 1u;
            }
        }
        if (t0 && !have_nxt) {  // (only when the look-ahead never reached the next strip: cannot happen with ahead >= 1)
            nxt = atomicAdd(p.ticket, 1u);
            s_ticket[(it + 1u) & 1u] = nxt;
        }
        have_nxt = false;
        bar_test_group();  // the next ticket and the next run's base are visible to the whole group
        cur = s_ticket[(it + 1u) & 1u];
    }
    // an entry that did not fit the staging buffer makes the result invalid: say so (the host maps it to an error)
    if (t0 && *s_dropped != 0u) atomicOr(p.flags, kFlagStagingOverflow);
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

// ---- ordered compaction, step 3: strip by strip, unordered runs -> row-major points ----------------------
// Persistent CTAs; a CTA takes strips blockIdx.x, blockIdx.x + gridDim.x, ...  For each strip its keypoints are
// scattered into a two-level bitmap in shared memory: level 1 = one bit per pixel of the strip's emitted rows
// (out_rows x words_per_row words), level 2 = one bit per level-1 word.  Thread t then owns level-1 words
// 32t .. 32t+31 (in row-major order): it counts their bits by walking the set bits of its level-2 word, a block
// prefix sum turns the counts into offsets, and the same walk expands the bits to (x, y) points at the strip's
// final position (fast_simd.rs:550, 596-613: the output is row-major).  Every word that is read is cleared, so
// the bitmap is zeroed only once per CTA and the work per strip is proportional to its keypoints.  The run
// records of the next strip are fetched while the current one is processed (the kernel is latency-bound).
struct StripRecord {
    unsigned long long dst;
    uint32_t total;
};

// own_scan != 0 (n_items <= kGatherThreads: one image or a handful): there was no scan launch; every CTA scans the
// strip counts itself (one per thread) and CTA 0 writes the CSR frame offsets -- one launch less on the latency path.
__global__ void __launch_bounds__(kGatherThreads) fdf_gather_kernel(const DetectParams p, uint32_t n_items,
                                                                    uint32_t own_scan) {
    extern __shared__ __align__(16) uint32_t gsm[];
    __shared__ uint32_t warp_sums[kGatherThreads / 32];
    __shared__ uint32_t s_item_dst[kGatherThreads];
    __shared__ unsigned long long s_run_base[kGatherMaxChunks];
    __shared__ uint32_t s_run_count[kGatherMaxChunks];
    __shared__ StripRecord s_rec;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int mode = (int)p.mode, sr = (int)p.sr;
    const int WW = (int)p.words_per_row, NC = (int)p.chunks_per_strip;
    const int nwords = out_rows(mode, sr) * WW;
    const int nsum = (nwords + 31) / 32;  // level-2 words
    uint32_t *bits = gsm, *summary = gsm + nsum * 32;
    for (int i = tid; i < nsum * 33; i += kGatherThreads) gsm[i] = 0u;
    // Programmatic dependent launch: this grid may have been started while the kernel in front of it (detection, or the
    // scan) was still running -- everything above touched only shared memory.  From here on that kernel's results are
    // complete and visible.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    // (the detection kernel is complete: its flags are final, and this kernel only repeats bit 2 where that one set it)
    if (blockIdx.x == 0 && tid == 0 && p.flags_copy != nullptr) *p.flags_copy = *p.flags;

    // records of a strip: thread c < NC holds chunk c's run, thread NC the strip's total and destination
    unsigned long long r_base = 0ull;
    uint32_t r_count = 0u;
    auto fetch = [&](uint32_t item) {
        r_base = 0ull;
        r_count = 0u;
        if (item >= n_items) return;
        if (tid < NC) {
            const size_t slot = (size_t)item * NC + tid;
            r_count = p.run_count[slot];
            r_base = r_count != 0u ? p.run_base[slot] : 0ull;
        } else if (tid == NC) {
            r_count = p.item_count[item];
            r_base = own_scan ? (unsigned long long)s_item_dst[item] : p.item_dst[item];
        }
    };
    if (own_scan) {
        const uint32_t v = (uint32_t)tid < n_items ? p.item_count[tid] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += u;
        }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        uint32_t before = 0u;
#pragma unroll
        for (int w2 = 0; w2 < kGatherThreads / 32; w2++)
            if (w2 < warp) before += warp_sums[w2];
        const uint32_t excl = before + incl - v;
        s_item_dst[tid] = excl;
        if (blockIdx.x == 0 && (uint32_t)tid < n_items) {
            if ((uint32_t)tid % p.strips_per_frame == 0u) p.offsets[(uint32_t)tid / p.strips_per_frame] = excl;
            if ((uint32_t)tid == n_items - 1u) p.offsets[p.n_frames] = (unsigned long long)excl + v;
        }
        __syncthreads();  // (also: warp_sums is free again)
    }
    fetch(blockIdx.x);
    // (row, column) of the first level-1 word of each round of this thread: fixed for the whole kernel
    const int row_first = (32 * tid) / WW, col_first = 32 * tid - row_first * WW;
    const int row_step = (32 * kGatherThreads) / WW, col_step = 32 * kGatherThreads - row_step * WW;
    for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        __syncthreads();  // the previous strip is finished with the records (and, the first time, the bitmap is zero)
        if (tid < NC) {
            s_run_base[tid] = r_base;
            s_run_count[tid] = r_count;
        } else if (tid == NC) {
            s_rec.dst = r_base;
            s_rec.total = r_count;
        }
        __syncthreads();
        fetch(item + gridDim.x);  // in flight while this strip is processed
        if (s_rec.total == 0u) continue;  // (block-uniform)
        const uint32_t strip = item % p.strips_per_frame;
        const uint32_t y0 = (uint32_t)(first_out_row(mode) + (int)strip * out_rows(mode, sr));
        // one warp per chunk: its run -> bits
        for (int c = warp; c < NC; c += kGatherThreads / 32) {
            const unsigned long long base = s_run_base[c];
            const uint32_t cnt = s_run_count[c];
            for (uint32_t i = (uint32_t)lane; i < cnt; i += 32u) {
                if (base + i >= p.staging_cap) {  // (the detection kernel that dropped these entries has raised the flag too)
                    atomicOr(p.flags, kFlagStagingOverflow);
                    break;
                }
                const uint32_t e = p.staging[base + i];
                const uint32_t x = e & 0xffffu, w1 = (e >> 16) * (uint32_t)WW + (x >> 5);
                atomicOr(&bits[w1], 1u << (x & 31u));
                atomicOr(&summary[w1 >> 5], 1u << (w1 & 31u));
            }
        }
        __syncthreads();
        const unsigned long long o = s_rec.dst;
        uint32_t block_off = 0u;
        int row0 = row_first, col0 = col_first;  // of level-1 word 32 * ts
        for (int t0 = 0; t0 < nsum; t0 += kGatherThreads) {  // (one round unless the image is wider than ~8000 pixels)
            const int ts = t0 + tid;
            // the level-2 word bit-reversed: its set bits are then walked from the top (one FLO each) in ascending word order
            const uint32_t sm = ts < nsum ? __brev(summary[ts]) : 0u;
            uint32_t cnt = 0u;
            for (uint32_t m = sm; m != 0u;) {
                const int b = __clz(m);
                m &= ~(0x80000000u >> b);
                cnt += (uint32_t)__popc(bits[32 * ts + b]);
            }
            uint32_t incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += v;
            }
            if (t0 != 0) __syncthreads();  // warp_sums of the previous round have been read
            if (lane == 31) warp_sums[warp] = incl;
            __syncthreads();
            uint32_t before = block_off, all = 0u;
#pragma unroll
            for (int w2 = 0; w2 < kGatherThreads / 32; w2++) {
                const uint32_t ws = warp_sums[w2];
                if (w2 < warp) before += ws;
                all += ws;
            }
            block_off += all;
            unsigned long long oo = o + before + (incl - cnt);
            for (uint32_t m = sm; m != 0u;) {
                const int b = __clz(m);
                m &= ~(0x80000000u >> b);
                const uint32_t word = bits[32 * ts + b];
                bits[32 * ts + b] = 0u;
                int row = row0, col = col0 + b;
                while (col >= WW) {
                    col -= WW;
                    row++;
                }
                emit_word(word, (uint32_t)col * 32u, y0 + (uint32_t)row, oo, p.cap, p.out);
                oo += (unsigned long long)__popc(word);
            }
            if (sm != 0u) summary[ts] = 0u;
            row0 += row_step;
            col0 += col_step;
            if (col0 >= WW) {
                col0 -= WW;
                row0++;
            }
        }
    }
}

// ---- sharded batches: this rank's points -> their place in the batch result (possibly another GPU's memory) ------
__global__ void __launch_bounds__(256) fdf_shard_push_kernel(const unsigned long long *all_offsets, uint32_t block,
                                                            uint32_t n_ranks, uint32_t rank, uint32_t total_frames,
                                                            const uint2 *points, uint2 *result,
                                                            unsigned long long cap_total,
                                                            unsigned long long *global_offsets) {
    unsigned long long base = 0ull, mine = 0ull;
    for (uint32_t r = 0; r <= rank; r++) {
        const uint32_t fr = shard_lo(total_frames, r + 1u, n_ranks) - shard_lo(total_frames, r, n_ranks);
        const unsigned long long tot = all_offsets[(size_t)r * block + fr];
        if (r < rank) base += tot;
        else mine = tot;
    }
    if (blockIdx.x == 0) {  // the batch's global CSR offsets: every rank's local offsets shifted by the lower ranks' totals
        unsigned long long rb = 0ull;
        for (uint32_t r = 0; r < n_ranks; r++) {
            const uint32_t lo = shard_lo(total_frames, r, n_ranks);
            const uint32_t fr = shard_lo(total_frames, r + 1u, n_ranks) - lo;
            const unsigned long long *blk = all_offsets + (size_t)r * block;
            for (uint32_t f = threadIdx.x; f < fr; f += blockDim.x) global_offsets[lo + f] = rb + blk[f];
            rb += blk[fr];
        }
        if (threadIdx.x == 0) global_offsets[total_frames] = rb;
    }
    if (result == nullptr || result + base == points) return;  // (nothing to move: the points already sit in place)
    if (base >= cap_total) return;
    if (base + mine > cap_total) mine = cap_total - base;  // (the caller sees offsets[total] > cap)
    uint2 *dst = result + base;
    const unsigned long long i0 = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long step = (unsigned long long)gridDim.x * blockDim.x;
    if (((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(points)) & 15u) == 0u) {
        const uint4 *s4 = reinterpret_cast<const uint4 *>(points);
        uint4 *d4 = reinterpret_cast<uint4 *>(dst);
        for (unsigned long long i = i0; i < mine / 2ull; i += step) d4[i] = s4[i];
        if ((mine & 1ull) && i0 == 0ull) dst[mine - 1ull] = points[mine - 1ull];
    } else {
        for (unsigned long long i = i0; i < mine; i += step) dst[i] = points[i];
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

// ---- RGB8 -> luma8: the step in front of the path (main.rs:53-58 `image::open(..).to_rgb8()` then `.to_luma8()`) ----
// image 0.24.6 (Cargo.lock:388-390, not vendored): luma = (2126 r + 7152 g + 722 b) / 10000 in u32, truncating
// (color.rs: SRGB_LUMA = [2126, 7152, 722], SRGB_LUMA_DIV = 10000), which is the identity for r == g == b.
// HBM-bound (3 bytes read, 1 written per pixel): one thread converts four pixels, 12 bytes in as three words when the
// row is word-aligned, one word out.
// kind 0: image 0.24.6 to_luma8 (above); kind 1: the crate's own util.rs:5-41 `Rgb8ToLuma16View` + `to_grey`:
// the view's pixel is r + g + b as u16, to_grey stores (that / 3) as u8.
template <int KIND>
__device__ __forceinline__ uint32_t luma_of(uint32_t r, uint32_t g, uint32_t b) {
    return KIND == 0 ? (2126u * r + 7152u * g + 722u * b) / 10000u : (r + g + b) / 3u;
}

template <int KIND>
__global__ void fdf_luma_kernel(const uint8_t *rgb, uint32_t n_frames, uint32_t w, uint32_t h, uint32_t rgb_pitch,
                                unsigned long long rgb_stride, uint8_t *luma, uint32_t luma_pitch,
                                unsigned long long luma_stride) {
    const unsigned long long w4 = (w + 3u) / 4u;
    const unsigned long long total = (unsigned long long)n_frames * h * w4;
    const bool aligned = ((reinterpret_cast<uintptr_t>(rgb) | rgb_pitch | rgb_stride) & 3u) == 0u &&
                         ((reinterpret_cast<uintptr_t>(luma) | luma_pitch | luma_stride) & 3u) == 0u;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const uint32_t x4 = (uint32_t)(i % w4);
        const uint32_t y = (uint32_t)((i / w4) % h);
        const uint32_t f = (uint32_t)(i / (w4 * h));
        const uint8_t *src = rgb + (size_t)f * rgb_stride + (size_t)y * rgb_pitch + (size_t)x4 * 12u;
        uint8_t *dst = luma + (size_t)f * luma_stride + (size_t)y * luma_pitch + (size_t)x4 * 4u;
        if (aligned && x4 * 4u + 4u <= w) {
            const uint32_t a = reinterpret_cast<const uint32_t *>(src)[0];  // r0 g0 b0 r1
            const uint32_t b = reinterpret_cast<const uint32_t *>(src)[1];  // g1 b1 r2 g2
            const uint32_t c = reinterpret_cast<const uint32_t *>(src)[2];  // b2 r3 g3 b3
            const uint32_t l0 = luma_of<KIND>(a & 0xffu, (a >> 8) & 0xffu, (a >> 16) & 0xffu);
            const uint32_t l1 = luma_of<KIND>(a >> 24, b & 0xffu, (b >> 8) & 0xffu);
            const uint32_t l2 = luma_of<KIND>((b >> 16) & 0xffu, b >> 24, c & 0xffu);
            const uint32_t l3 = luma_of<KIND>((c >> 8) & 0xffu, (c >> 16) & 0xffu, c >> 24);
            *reinterpret_cast<uint32_t *>(dst) = l0 | (l1 << 8) | (l2 << 16) | (l3 << 24);
        } else {
            for (uint32_t k = 0; k < 4u && x4 * 4u + k < w; k++)
                dst[k] = (uint8_t)luma_of<KIND>(src[3 * k], src[3 * k + 1], src[3 * k + 2]);
        }
    }
}

#ifdef FDF_CHECKS
}  // namespace
cudaError_t read_check_failure(int *line) {  // and resets it
    cudaError_t e = cudaMemcpyFromSymbol(line, g_check_failure_line, sizeof(int));
    if (e != cudaSuccess) return e;
    const int zero = 0;
    return cudaMemcpyToSymbol(g_check_failure_line, &zero, sizeof(zero));
}
namespace {
#endif

#ifdef FDF_TRACE
}  // namespace
cudaError_t read_trace(long long *out, size_t bytes) {  // and resets the CTA counter for the next launch
    if (bytes > sizeof(g_trace)) bytes = sizeof(g_trace);
    cudaError_t e = cudaMemcpyFromSymbol(out, g_trace, bytes);
    if (e != cudaSuccess) return e;
    const unsigned int zero = 0u;
    return cudaMemcpyToSymbol(g_trace_n, &zero, sizeof(zero));
}
namespace {
#endif

template <int MODE, int SR>
cudaError_t prepare_t(int *per_sm) {
    auto kern = fdf_detect_kernel<MODE, SR>;
    const size_t smem = detect_smem_bytes(MODE, SR);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, kern, kThreads, smem);
}

template <int MODE, int SR>
cudaError_t launch_t(const CUtensorMap &tmap, const DetectParams &p, cudaStream_t stream, unsigned grid) {
    fdf_detect_kernel<MODE, SR><<<grid, kThreads, detect_smem_bytes(MODE, SR), stream>>>(tmap, p);
    return cudaGetLastError();
}

constexpr size_t kGatherSmemLimit = 200 * 1024;

}  // namespace

size_t detect_smem_bytes(int mode, int sr) {
    (void)mode;
    return (size_t)layout_sizes(sr).total;
}

size_t gather_smem_bytes(int mode, int sr, uint32_t words_per_row) {
    const size_t nsum = ((size_t)out_rows(mode, sr) * words_per_row + 31) / 32;  // level-2 words
    return nsum * 33 * 4;
}

// Everything a launch needs to know about the device and the kernels is looked up ONCE per context (fdf_create):
// function attributes, occupancy per (mode, strip height), SM count, the experiment knob.  A detection call then
// costs three launches and nothing else on the host.
cudaError_t init_device_info(DeviceInfo &info) {
    int dev = 0;
    cudaError_t e;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&info.sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
#define FDF_CASE(M, S, I) \
    if ((e = prepare_t<M, S>(&info.detect_per_sm[M][I])) != cudaSuccess) return e;
    FDF_CASE(0, 32, 0) FDF_CASE(0, 48, 1) FDF_CASE(0, 64, 2)
    FDF_CASE(1, 32, 0) FDF_CASE(1, 48, 1) FDF_CASE(1, 64, 2)
    FDF_CASE(2, 32, 0) FDF_CASE(2, 48, 1) FDF_CASE(2, 64, 2)
#undef FDF_CASE
    e = cudaFuncSetAttribute(fdf_gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGatherSmemLimit);
    if (e != cudaSuccess) return e;
    info.gather_smem = ~(size_t)0;
    info.gather_per_sm = 0;
    info.ctas_limit = 0;
    if (const char *lim = getenv("FDF_CTAS_PER_SM")) info.ctas_limit = atoi(lim);  // tuning knob for experiments
    return cudaSuccess;
}

cudaError_t launch_detect(int mode, int sr, const CUtensorMap &tmap, const DetectParams &p, cudaStream_t stream,
                          const DeviceInfo &info) {
    const unsigned long long items = (unsigned long long)p.n_frames * p.strips_per_frame;
    if (items == 0 || items > 0x7fffffffull || mode < 0 || mode > 2) return cudaErrorInvalidValue;
    const int si = sr == 32 ? 0 : (sr == 48 ? 1 : (sr == 64 ? 2 : -1));
    if (si < 0) return cudaErrorInvalidValue;
    int per_sm = info.detect_per_sm[mode][si];
    if (per_sm < 1) return cudaErrorInvalidConfiguration;
    if (info.ctas_limit >= 1 && info.ctas_limit < per_sm) per_sm = info.ctas_limit;
    // persistent grid: as many CTAs as can be resident at once (each loops over tickets)
    unsigned long long grid = (unsigned long long)info.sms * (unsigned)per_sm;
    const unsigned long long tickets = items * (sr == 32 && p.parts > 1u ? p.parts : 1u);  // (only the 32-row kernels split)
    if (grid > tickets) grid = tickets;
#define FDF_CASE(M, S) \
    if (mode == M && sr == S) return launch_t<M, S>(tmap, p, stream, (unsigned)grid);
    FDF_CASE(0, 32)
    FDF_CASE(0, 48)
    FDF_CASE(0, 64)
    FDF_CASE(1, 32)
    FDF_CASE(1, 48)
    FDF_CASE(1, 64)
    FDF_CASE(2, 32)
    FDF_CASE(2, 48)
    FDF_CASE(2, 64)
#undef FDF_CASE
    return cudaErrorInvalidValue;
}

bool gather_scans_itself(const DetectParams &p) {
    return (unsigned long long)p.n_frames * p.strips_per_frame <= (unsigned long long)kGatherThreads;
}

cudaError_t launch_scan(const DetectParams &p, cudaStream_t stream) {
    const unsigned long long items = (unsigned long long)p.n_frames * p.strips_per_frame;
    if (items == 0 || items > 0x7fffffffull) return cudaErrorInvalidValue;
    const unsigned tiles = (unsigned)((items + kScanTile - 1) / kScanTile);
    fdf_scan_kernel<<<tiles, kScanThreads, 0, stream>>>(p, (uint32_t)items);
    return cudaGetLastError();
}

cudaError_t launch_gather(const DetectParams &p, cudaStream_t stream, DeviceInfo &info) {
    const unsigned long long items = (unsigned long long)p.n_frames * p.strips_per_frame;
    if (items == 0 || items > 0x7fffffffull) return cudaErrorInvalidValue;
    const size_t smem = gather_smem_bytes((int)p.mode, (int)p.sr, p.words_per_row);
    if (smem > kGatherSmemLimit) return cudaErrorInvalidValue;
    if (smem != info.gather_smem) {  // (occupancy depends on the image width only: looked up when the width changes)
        int per_sm = 0;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fdf_gather_kernel, kGatherThreads, smem);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) return cudaErrorInvalidConfiguration;
        info.gather_smem = smem;
        info.gather_per_sm = per_sm;
    }
    unsigned long long grid = (unsigned long long)info.sms * (unsigned)info.gather_per_sm;
    if (grid > items) grid = items;
    // launched with programmatic stream serialisation: its CTAs may be placed and run their prologue (bitmap clear)
    // while the kernel in front drains; griddepcontrol.wait in the kernel keeps the data dependence (one launch gap
    // less on the single-image latency path)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)kGatherThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, fdf_gather_kernel, p, (uint32_t)items, gather_scans_itself(p) ? 1u : 0u);
}

cudaError_t launch_shard_push(const unsigned long long *all_offsets, uint32_t block, uint32_t n_ranks, uint32_t rank,
                              uint32_t total_frames, const uint2 *points, uint2 *result, unsigned long long cap_total,
                              unsigned long long *global_offsets, int sms, cudaStream_t stream) {
    // a fraction of the SMs is enough to fill NVLink; the rest stays free for the next step's detection kernel
    const unsigned grid = (unsigned)(sms > 0 ? (sms + 3) / 4 : 32);
    fdf_shard_push_kernel<<<grid, 256, 0, stream>>>(all_offsets, block, n_ranks, rank, total_frames, points, result,
                                                   cap_total, global_offsets);
    return cudaGetLastError();
}

cudaError_t launch_luma(const uint8_t *d_rgb, uint32_t n_frames, uint32_t w, uint32_t h, uint32_t rgb_pitch,
                        unsigned long long rgb_stride, uint8_t *d_luma, uint32_t luma_pitch,
                        unsigned long long luma_stride, int kind, cudaStream_t stream) {
    const unsigned long long total = (unsigned long long)n_frames * h * ((w + 3u) / 4u);
    if (total == 0) return cudaSuccess;
    unsigned long long blocks = (total + 255ull) / 256ull;
    if (blocks > 148ull * 32ull) blocks = 148ull * 32ull;  // grid-stride: a multiple of the SM count
    if (kind == 0)
        fdf_luma_kernel<0><<<(unsigned)blocks, 256, 0, stream>>>(d_rgb, n_frames, w, h, rgb_pitch, rgb_stride, d_luma,
                                                                 luma_pitch, luma_stride);
    else
        fdf_luma_kernel<1><<<(unsigned)blocks, 256, 0, stream>>>(d_rgb, n_frames, w, h, rgb_pitch, rgb_stride, d_luma,
                                                                 luma_pitch, luma_stride);
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
