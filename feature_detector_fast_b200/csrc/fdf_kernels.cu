// fdf_kernels.cu -- sm_100a kernels of the FAST-n detection path.
//
// One persistent kernel does the detection for a batch of frames (replaces fast_simd.rs:301-620):
//
//   work item  = (frame, strip of full-width rows), handed out through an atomic ticket.
//   8 warps, per chunk of a strip (two block barriers per chunk):
//     TMA 3-D tiled load (u8 tile 256 x (SR+6), zero-filled outside the image, double buffered,
//     issued two chunks ahead -- across strip boundaries --, completion on an mbarrier)  -> shared memory
//     phase A  : two-stage dense SWAR filter.  Stage 1 (every pixel, 16 per lane: LDS.128 + VABSDIFF4 + LOP3)
//                tests the north/south pair; groups with a survivor go to the warp's own queue (ballot, no
//                barrier); stage 2 (one lane per queued group) adds the east/west pair (PRMT byte shifts) and
//                pushes the surviving centres to the CTA's candidate queue           (fast_simd.rs:368-520)
//                (the NMS pass of the previous chunk runs in the same barrier interval)
//     phase B  : one thread per candidate: 16 ring bytes, two per 32-bit word -> brighter/darker 16-bit masks
//                -> rotate-AND arc test -> score in 16-bit lanes -> tagged score plane + keypoint list
//                                                                     (fast_simd.rs:115-297, 623-749)
//                (one warp meanwhile copies the previous chunk's surviving keypoints to the staging buffer)
//     NMS pass : strict 3x3 maximum on the shared-memory score plane          (fast_simd.rs:588-616)
//                survivors are marked in the chunk's keypoint list
//   Every chunk leaves one unordered run of (row, x) entries in the staging buffer plus a run record.
//
// Two small kernels finish the ordered compaction (fast_simd.rs:550, 596-613: output is row-major):
//   fdf_scan_kernel   : exclusive prefix sum of the per-strip counts in (frame, strip) order -- block scan
//                       + decoupled look-back between scan tiles -- giving every strip's final offset and
//                       the CSR frame offsets;
//   fdf_gather_kernel : one CTA per strip scatters the strip's runs into a bit plane in shared memory and
//                       expands it, row-major, to (x, y) points at the strip's final offset.
// (Doing the look-back inside the detection kernel was measured at +45 % kernel time: with ~450 strips
// in flight every strip ends up waiting for all in-flight predecessors.  Keeping the strip bit plane inside
// the detection kernel cost 30 KB of shared memory per CTA, i.e. one resident CTA per SM.)
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
#ifndef FDF_WAIT_MODE
#define FDF_WAIT_MODE 0
#endif
#ifndef FDF_WAIT_HINT_NS
#define FDF_WAIT_HINT_NS 100000
#endif
constexpr uint32_t kWaitHintNs = FDF_WAIT_HINT_NS;  // mbarrier.try_wait suspend-time hint
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

// Spins on mbarrier.try_wait (each probe suspends the thread in hardware until the phase completes or the time hint
// runs out) for at most `rounds` probes: the loop is three instructions -- probe, branch out, count-and-branch back --
// because waiting warps re-issue it all the time and every extra instruction in it is an issue slot (and a logic-pipe
// slot) taken from the warps that work.  Returns true when the phase has completed.
__device__ __forceinline__ bool mbar_spin(uint64_t *bar, uint32_t parity, uint32_t rounds) {
    uint32_t done;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .u32 n;\n"
        "mov.u32 n, %3;\n"
        "FDF_SPIN:\n"
#if FDF_WAIT_MODE == 1   // experiment: no suspend-time hint (the hardware's default time limit)
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
#elif FDF_WAIT_MODE == 2  // experiment: non-blocking test (busy polling)
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
#else
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %4;\n"
#endif
        "@p bra FDF_SPIN_DONE;\n"
        "add.u32 n, n, -1;\n"
        "setp.ne.u32 p, n, 0;\n"
        "@p bra FDF_SPIN;\n"
        "setp.ne.u32 p, n, n;\n"  // (gave up: p = false)
        "FDF_SPIN_DONE:\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity), "r"(rounds), "r"(kWaitHintNs)
        : "memory");
    return done != 0;
}

// Waits for the phase; a wait that lasts longer than kWaitLimitNs can only be a bug in the pipeline: it is turned
// into an error flag (the host reports FDF_ERR_INTERNAL) instead of a hung GPU.  The clock and the CTA's abort flag
// are looked at once per kSpinRounds probes only.
// (`abort` is a flag in shared memory: once one wait of the CTA has timed out, no other wait of the CTA blocks, so
// that the kernel still ends quickly.)
constexpr uint32_t kSpinRounds = 4096u;
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity, uint32_t *flags, volatile uint32_t *abort) {
    if (mbar_spin(bar, parity, kSpinRounds)) return;
    const unsigned long long t0 = global_timer_ns();
    while (!mbar_spin(bar, parity, kSpinRounds)) {
        if (*abort != 0u) return;
        if (global_timer_ns() - t0 > kWaitLimitNs) {
            atomicOr(flags, kFlagTmaTimeout);
            *abort = 1u;
            return;
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
                                              // the test warps are done with chunk k
constexpr int kQueueBufs = kTileStages;       // candidate queues, one per tile buffer
constexpr int kTicketSlots = 8;               // strips whose tickets can be in flight between the tile requests and the
                                              // emit warps (a strip can be a single chunk)
constexpr uint32_t kDenseMark = 0xffffffffu;  // keypoint count of a chunk that went through the dense path
static_assert(kTileStages >= 2 && kTileStages <= 4, "misc layout holds up to 4 barriers of each kind");

// score planes / keypoint lists per CTA: two one-byte planes let the test warps fill the plane of chunk k + 1 while
// the emit warps still read the plane of chunk k; SumAbsolute needs two-byte cells, so it gets one plane and the two
// groups take turns on it
__host__ __device__ constexpr int plane_bufs(int mode) { return mode == NMS_SUM_ABSOLUTE ? 1 : 2; }
__host__ __device__ constexpr int plane_cell_bytes(int mode) { return mode == NMS_SUM_ABSOLUTE ? 2 : 1; }

struct LayoutSizes {
    int tile_bytes, plane_off, plane_bytes, queue_off, klist_off, ent_off, vtab_off, misc_off, total;
};
__host__ __device__ constexpr LayoutSizes layout_sizes(int mode, int sr) {
    LayoutSizes l = {};
    l.tile_bytes = tile_rows(sr) * kTileW;                                  // one TMA box
    l.plane_off = kTileStages * l.tile_bytes;
    l.plane_bytes = sr * kPlaneW * plane_cell_bytes(mode);                  // one plane
    l.queue_off = l.plane_off + plane_bufs(mode) * l.plane_bytes;           // candidate queues [kQueueBufs][kQueueCap] u16
    l.klist_off = l.queue_off + kQueueBufs * kQueueCap * 2;                 // keypoint lists [planes][kKlistCap] u16
    l.ent_off = l.klist_off + plane_bufs(mode) * kKlistCap * 2;             // stage-1 entry tables [2][warps][kWarpQueueCap] u8
    l.vtab_off = l.ent_off + 2 * kFilterWarps * kWarpQueueCap;              // validity tables: first / middle / last chunk
    l.misc_off = l.vtab_off + 3 * kVtabWords * 4;
    l.total = l.misc_off + 256;
    return l;
}

// misc block (byte offsets)
constexpr int kMiscFullBar = 0;     // [4] u64 tile landed
constexpr int kMiscQFull = 32;      // [4] u64 candidate queue complete (one arrival per filter warp)
constexpr int kMiscKFull = 64;      // [2] u64 score plane + keypoint list complete (the last test warp arrives)
constexpr int kMiscPFree = 80;      // [2] u64 score plane + keypoint list free again (the last emit warp arrives)
constexpr int kMiscQCount = 96;     // [4] u32 candidate queue fill
constexpr int kMiscKCount = 112;    // [2] u32 keypoints of the chunk
constexpr int kMiscTDone = 120;     // [4] u32 test warps done with the chunk
constexpr int kMiscEDone = 136;     // [2] u32 emit warps done with the chunk
constexpr int kMiscTicket = 144;    // [8] u32 strip tickets
constexpr int kMiscNEnt = 176;      // [2][4] u32 stage-1 entries per filter warp
constexpr int kMiscReqDone = 208;   // u32 tile requests issued so far (they are issued in stream order)
constexpr int kMiscAbort = 212;     // u32 a wait timed out
constexpr int kMiscTrace = 216;     // i32 (trace builds) index of this CTA in the trace table
static_assert(kFilterWarps <= 4 && kTicketSlots == 8, "misc layout");

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
// Persistent CTAs; a CTA draws strips (frame, rows) from an atomic ticket and walks each strip chunk by chunk.  Its
// warps form a three-stage pipeline over the stream of chunks, coupled ONLY through mbarriers and a few shared
// counters -- no CTA-wide barrier, and no barrier at all inside the test and emit groups, whose warps drift freely:
//
//   filter warps (kFilterWarps)   wait for the chunk's tile (TMA, full_bar) -> stage 1 over their own rows -> one
//                                 barrier among the filter warps -> stage 2 over ALL warps' entries, dealt evenly ->
//                                 candidate queue (one per tile buffer) -> arrive on q_full.
//   test warps (kTestWarps)       wait q_full and p_free (the plane and keypoint list of chunk k - planes are free) ->
//                                 exact test + score per candidate -> score plane, keypoint list.  The LAST test warp
//                                 to finish a chunk (shared counter) arrives on k_full and requests the tile of chunk
//                                 k + kTileStages: the tile and the candidate queue of chunk k are free now.
//   emit warps (kEmitWarps)       wait k_full -> strict 3x3 maximum per listed keypoint on the plane -> survivors into
//                                 the warp's own run of the staging buffer (ballot-ranked, no atomics), run record.  The
//                                 LAST emit warp wipes the chunk's cells from the plane and arrives on p_free.
//
// Why this shape (profiles/r02_v15_timeline_3ctas.txt): with per-warp stage-2 queues one filter warp regularly took
// 3-4 times as long as the others (a horizontal edge puts most of a chunk's 16-pixel groups into one warp's rows),
// the test group waited for it, the filter group then waited for the test group's tile request, and the test group's
// own chain (test -> barrier -> tile request -> NMS -> barrier -> bookkeeping by one thread) was the longest stage.
// Stream bookkeeping: a chunk is (strip sequence number `it` of this CTA, chunk c of the strip); every warp walks the
// same sequence on its own.  The ticket of strip `it` sits in s_ticket[it % kTicketSlots]; it is drawn by the thread
// that requests the strip's first tile.  Tile requests are issued in stream order (s_req_done), so tickets are drawn in
// order and the first ticket beyond the last strip ends the stream for everybody.
__device__ __forceinline__ void bar_filter_group() {
    asm volatile("bar.sync 1, %0;" ::"n"(kFilterThreads) : "memory");
}

__device__ __forceinline__ uint32_t ld_volatile_shared(const uint32_t *p) {
    return *reinterpret_cast<const volatile uint32_t *>(p);
}

template <int MODE, int SR>
__global__ void __launch_bounds__(kThreads, SR >= 48 ? 3 : 4)
fdf_detect_kernel(const __grid_constant__ CUtensorMap tmap, const DetectParams p) {
    typedef typename PlaneCell<MODE>::type cell_t;
    constexpr LayoutSizes L = layout_sizes(MODE, SR);
    constexpr int PL = plane_bufs(MODE);
    constexpr int HS = (MODE == NMS_OFF) ? 0 : 1;  // score halo (rows and columns) needed by the 3x3 NMS
    constexpr int OUT_R = out_rows(MODE, SR);
    static_assert(L.tile_bytes % 128 == 0, "TMA destination must stay 128-byte aligned");
    static_assert(L.plane_bytes % 16 == 0 && L.plane_off % 16 == 0, "the plane is cleared with 128-bit stores");
    static_assert(SR % (2 * kFilterWarps) == 0 && SR <= 64, "filter warps take row pairs; queue entries hold 6 row bits");
    static_assert(L.vtab_off % 16 == 0 && L.misc_off % 8 == 0, "alignment of the validity tables / mbarriers");

    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *tiles = smem;
    cell_t *planes = reinterpret_cast<cell_t *>(smem + L.plane_off);                     // [PL][SR * kPlaneW]
    uint16_t *queues = reinterpret_cast<uint16_t *>(smem + L.queue_off);                 // [kQueueBufs][kQueueCap]
    uint16_t *klists = reinterpret_cast<uint16_t *>(smem + L.klist_off);                 // [PL][kKlistCap]
    uint8_t *ents = smem + L.ent_off;                                                    // [2][kFilterWarps][kWarpQueueCap]
    uint32_t *vtabs = reinterpret_cast<uint32_t *>(smem + L.vtab_off);                   // [3][kVtabWords]
    uint8_t *misc = smem + L.misc_off;
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(misc + kMiscFullBar);
    uint64_t *q_full = reinterpret_cast<uint64_t *>(misc + kMiscQFull);
    uint64_t *k_full = reinterpret_cast<uint64_t *>(misc + kMiscKFull);
    uint64_t *p_free = reinterpret_cast<uint64_t *>(misc + kMiscPFree);
    uint32_t *qcount = reinterpret_cast<uint32_t *>(misc + kMiscQCount);
    uint32_t *kcount = reinterpret_cast<uint32_t *>(misc + kMiscKCount);
    uint32_t *t_done = reinterpret_cast<uint32_t *>(misc + kMiscTDone);
    uint32_t *e_done = reinterpret_cast<uint32_t *>(misc + kMiscEDone);
    uint32_t *s_ticket = reinterpret_cast<uint32_t *>(misc + kMiscTicket);
    uint32_t *s_nent = reinterpret_cast<uint32_t *>(misc + kMiscNEnt);
    uint32_t *s_req_done = reinterpret_cast<uint32_t *>(misc + kMiscReqDone);
    volatile uint32_t *s_abort = reinterpret_cast<volatile uint32_t *>(misc + kMiscAbort);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int W = (int)p.w, H = (int)p.h;
    const int NC = (int)p.chunks_per_strip;
    const uint32_t total_items = p.n_frames * p.strips_per_frame;

    // Requests the tile of stream position r = chunk cr of strip sequence number itr (one thread).  Requests are issued
    // in stream order: the caller for position r waits until the requests 0 .. r-1 have been issued, which also makes
    // the ticket of a strip, drawn with its first tile, visible to the requests of its other tiles.  When the tickets
    // are exhausted the barrier is completed without a tile, so that the filter warps wake up and see the end ticket.
    auto request_tile = [&](uint32_t r, uint32_t itr, int cr) {
        const unsigned long long t_begin = global_timer_ns();
        while (ld_volatile_shared(s_req_done) != r) {
            if (*s_abort != 0u) return;
            if (global_timer_ns() - t_begin > kWaitLimitNs) {
                atomicOr(p.flags, kFlagTmaTimeout);
                *s_abort = 1u;
                return;
            }
        }
        __threadfence_block();
        uint32_t item;
        if (cr == 0) {
            item = atomicAdd(p.ticket, 1u);
            s_ticket[itr % (uint32_t)kTicketSlots] = item;
        } else {
            item = s_ticket[itr % (uint32_t)kTicketSlots];
        }
        const uint32_t stage = r % (uint32_t)kTileStages;
        if (item >= total_items) {
            mbar_arrive(&full_bar[stage]);
        } else {
            const uint32_t frame = item / p.strips_per_frame;
            const uint32_t strip = item - frame * p.strips_per_frame;
            const int ty0 = first_out_row(MODE) + (int)strip * OUT_R - HS - 3;  // image row of tile row 0
            mbar_expect_tx(&full_bar[stage], (uint32_t)L.tile_bytes);
            tma_load_3d(tiles + stage * L.tile_bytes, &tmap, cr * kChunkW - kTileLead, ty0, (int)frame, &full_bar[stage]);
            // and pull the strip's next tile into L2 (cp.async.bulk.prefetch.tensor: no shared memory, no completion), so
            // that its load, one chunk from now, is an L2 hit
            if (FDF_L2_PREFETCH > 0 && cr + FDF_L2_PREFETCH < NC)
                tma_prefetch_l2_3d(&tmap, (cr + FDF_L2_PREFETCH) * kChunkW - kTileLead, ty0, (int)frame);
        }
        __threadfence_block();
        *reinterpret_cast<volatile uint32_t *>(s_req_done) = r + 1u;
    };

    if (tid == 0) {
        tma_prefetch_desc(&tmap);
        for (int i = 0; i < kTileStages; i++) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&q_full[i], kFilterWarps);
            qcount[i] = 0u;
            t_done[i] = 0u;
        }
        for (int i = 0; i < 2; i++) {
            mbar_init(&k_full[i], 1);
            mbar_init(&p_free[i], 1);
            kcount[i] = 0u;
            e_done[i] = 0u;
        }
        fence_mbar_init();
        *s_req_done = 0u;
        *s_abort = 0u;
    }
    if (tid < 3 * kVtabWords) vtabs[tid] = valid_word<MODE>(W, vtab_chunk(tid / kVtabWords, NC), tid % kVtabWords);
    {
        uint4 *pz = reinterpret_cast<uint4 *>(planes);
        for (int i = tid; i < PL * L.plane_bytes / 16; i += kThreads) pz[i] = make_uint4(0u, 0u, 0u, 0u);
    }
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
    if (tid == 0) {  // the first kTileStages tiles of the stream
        uint32_t itr = 0u;
        int cr = 0;
        for (uint32_t r = 0; r < (uint32_t)kTileStages; r++) {
            request_tile(r, itr, cr);
            if (++cr == NC) {
                cr = 0;
                itr++;
            }
        }
    }

    const int t = (int)p.threshold, n = (int)p.count;
    uint32_t gc = 0;            // stream position of the chunk this warp works on
    uint32_t qb = 0, qpar = 0;  // tile stage = queue buffer = gc % kTileStages, and the parity of their mbarrier phases
    uint32_t pi = 0, ppar = 0;  // plane / keypoint list = gc % PL, and the parity of their mbarrier phases
    uint32_t cur = 0;           // ticket (frame, strip) of the current strip
    auto next_chunk = [&]() {
        gc++;
        if (++qb == (uint32_t)kQueueBufs) {
            qb = 0;
            qpar ^= 1u;
        }
        if (++pi == (uint32_t)PL) {
            pi = 0;
            ppar ^= 1u;
        }
    };

    if (warp < kFilterWarps) {
        // ================================ filter warps =================================================
        const uint32_t kbias = filter_kbias(p.threshold);
        for (uint32_t it = 0;; it++) {
            for (int c = 0; c < NC; c++) {
                mbar_wait(&full_bar[qb], qpar, p.flags, s_abort);  // the tile has landed (also: queue qb is free again)
                if (c == 0) {
                    cur = s_ticket[it % (uint32_t)kTicketSlots];
                    if (cur >= total_items || *s_abort != 0u) {  // the stream has ended: pass the news on and leave
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&q_full[qb]);
                        return;
                    }
                }
                FDF_CLK(0)
                const uint32_t frame = cur / p.strips_per_frame;
                const uint32_t strip = cur - frame * p.strips_per_frame;
                const ChunkGeo g = make_geo<MODE>(W, H, (int)strip, c, SR);
                const uint8_t *tile = tiles + qb * L.tile_bytes;
                uint8_t *eb = ents + (gc & 1u) * (kFilterWarps * kWarpQueueCap);
                uint32_t *nent = s_nent + (gc & 1u) * 4u;
                const uint32_t ne = phase_a_stage1<MODE, SR, kFilterWarps>(warp, lane, tile, eb + warp * kWarpQueueCap, g, kbias);
                if (lane == 0) nent[warp] = ne;
                FDF_CLK(9)
                bar_filter_group();  // every warp's entries of this chunk are in the table
                FDF_CLK(10)
                phase_a_stage2<MODE, SR, kFilterWarps>(tid, kFilterThreads, tile, eb, nent,
                                                       vtabs + vtab_variant(c, NC) * kVtabWords, queues + qb * kQueueCap,
                                                       &qcount[qb], kbias);
                __syncwarp();
                if (lane == 0) mbar_arrive(&q_full[qb]);
                FDF_CLK(1)
                next_chunk();
            }
        }
    } else if (warp < kFilterWarps + kTestWarps) {
        // ==================================== test warps ===================================================
        const int twarp = warp - kFilterWarps, ttid = tid - kFilterThreads;
        for (uint32_t it = 0;; it++) {
            for (int c = 0; c < NC; c++) {
                mbar_wait(&q_full[qb], qpar, p.flags, s_abort);
                mbar_wait(&full_bar[qb], qpar, p.flags, s_abort);  // (completed long ago: makes the tile visible here too)
                if (c == 0) cur = s_ticket[it % (uint32_t)kTicketSlots];
                const bool ended = cur >= total_items || *s_abort != 0u;
                mbar_wait(&p_free[pi], ppar ^ 1u, p.flags, s_abort);  // plane pi and its keypoint list are free
                FDF_CLK(4)
                uint32_t qn = 0u;
                if (!ended) {
                    const uint32_t frame = cur / p.strips_per_frame;
                    const uint32_t strip = cur - frame * p.strips_per_frame;
                    const uint8_t *tile = tiles + qb * L.tile_bytes;
                    cell_t *plane = planes + pi * (SR * kPlaneW);
                    qn = qcount[qb];
                    if (qn <= (uint32_t)kQueueCap) {
                        phase_b<MODE, SR>(ttid, lane, kTestThreads, qn, tile, queues + qb * kQueueCap, klists + pi * kKlistCap,
                                          &kcount[pi], plane, t, n);
                    } else {  // very dense content: every pixel, straight from the tile
                        const ChunkGeo g = make_geo<MODE>(W, H, (int)strip, c, SR);
                        phase_b_dense<MODE, SR>(twarp, lane, kTestWarps, tile, plane, g, c, t, n);
                    }
                }
                FDF_CLK(5)
                __syncwarp();
                // the last test warp to get here hands the chunk on and recycles its tile and candidate queue
                uint32_t last = 0u;
                if (lane == 0) {
                    __threadfence_block();
                    last = atomicAdd(&t_done[qb], 1u) == (uint32_t)(kTestWarps - 1) ? 1u : 0u;
                    if (last) {
                        __threadfence_block();
                        t_done[qb] = 0u;
                        if (!ended) {
                            if (qn > (uint32_t)kQueueCap) kcount[pi] = kDenseMark;
                            qcount[qb] = 0u;
                        }
                        __threadfence_block();
                        mbar_arrive(&k_full[pi]);
                        if (!ended) {
                            uint32_t itr = it;
                            int cr = c + kTileStages;
                            while (cr >= NC) {
                                cr -= NC;
                                itr++;
                            }
                            request_tile(gc + (uint32_t)kTileStages, itr, cr);
                        }
                    }
                }
                FDF_CLK(6)
                if (ended) return;
                next_chunk();
            }
        }
    } else {
        // ==================================== emit warps ===================================================
        const int ewarp = warp - kFilterWarps - kTestWarps;
        constexpr int kEmitThreads = kEmitWarps * 32;
        const uint32_t lt = (1u << lane) - 1u;
        unsigned long long blk_next = 0ull, blk_end = 0ull;  // this warp's staging block (warp-uniform)
        uint32_t strip_total = 0u;                           // keypoints this warp staged for the current strip
        bool overflow = false;
        // room for `need` entries in the warp's block, or a new block from the global cursor
        auto ensure_room = [&](uint32_t need) {
            if (blk_next + need > blk_end) {
                const unsigned long long nblk = need > (uint32_t)kStageBlock ? need : (unsigned long long)kStageBlock;
                unsigned long long b = 0ull;
                if (lane == 0) b = atomicAdd(p.cursor, nblk);
                blk_next = __shfl_sync(0xffffffffu, b, 0);
                blk_end = blk_next + nblk;
            }
        };
        for (uint32_t it = 0;; it++) {
            for (int c = 0; c < NC; c++) {
                mbar_wait(&k_full[pi], ppar, p.flags, s_abort);
                if (c == 0) cur = s_ticket[it % (uint32_t)kTicketSlots];
                if (cur >= total_items || *s_abort != 0u) {
                    if (overflow && lane == 0) atomicOr(p.flags, kFlagStagingOverflow);
                    return;
                }
                FDF_CLK(7)
                const uint32_t frame = cur / p.strips_per_frame;
                const uint32_t strip = cur - frame * p.strips_per_frame;
                const ChunkGeo g = make_geo<MODE>(W, H, (int)strip, c, SR);
                const cell_t *plane = planes + pi * (SR * kPlaneW);
                const uint16_t *klist = klists + pi * kKlistCap;
                const uint32_t kn = kcount[pi];
                uint32_t count = 0u;
                if (kn <= (uint32_t)kKlistCap) {
                    // this warp's share of the keypoint list: entries ewarp * 32 + lane, + kEmitThreads, ...
                    ensure_room(kn);
                    for (uint32_t ib = (uint32_t)(ewarp * 32); ib < kn; ib += (uint32_t)kEmitThreads) {
                        const uint32_t i = ib + (uint32_t)lane;
                        const uint32_t ent = klist[i < kn ? i : ib];
                        const bool keep = i < kn && list_entry_survives<MODE, SR>(ent, plane, g);
                        const uint32_t b = __ballot_sync(0xffffffffu, keep);
                        if (keep) {
                            const unsigned long long o = blk_next + count + (uint32_t)__popc(b & lt);
                            if (o < p.staging_cap) p.staging[o] = staged_entry<MODE>((int)(ent >> 8), (int)(ent & 0xffu), g);
                            else overflow = true;
                        }
                        count += (uint32_t)__popc(b);
                    }
                } else {
                    // the list overflowed or the chunk went through the dense path: scan this warp's rows of the plane,
                    // count first (the run is reserved at its exact size), then write
                    constexpr int kCells = SR * kPlaneW, kPer = (kCells + kEmitWarps - 1) / kEmitWarps;
                    const int i0 = ewarp * kPer, i1 = min(kCells, i0 + kPer);
                    for (int ib = i0; ib < i1; ib += 32) {
                        const int i = ib + lane;
                        const bool keep = i < i1 && plane_cell_survives<MODE, SR>(i, plane, g);
                        count += (uint32_t)__popc(__ballot_sync(0xffffffffu, keep));
                    }
                    ensure_room(count);
                    uint32_t at = 0u;
                    for (int ib = i0; ib < i1; ib += 32) {
                        const int i = ib + lane;
                        const bool keep = i < i1 && plane_cell_survives<MODE, SR>(i, plane, g);
                        const uint32_t b = __ballot_sync(0xffffffffu, keep);
                        if (keep) {
                            const unsigned long long o = blk_next + at + (uint32_t)__popc(b & lt);
                            if (o < p.staging_cap) p.staging[o] = staged_entry<MODE>(i / kPlaneW, i % kPlaneW + kPlaneLead, g);
                            else overflow = true;
                        }
                        at += (uint32_t)__popc(b);
                    }
                }
                FDF_CLK(8)
                if (lane == 0) {  // the warp's run record of this chunk
                    const size_t slot = ((size_t)cur * (size_t)NC + (size_t)c) * kRunsPerChunk + (size_t)ewarp;
                    p.run_base[slot] = blk_next;
                    p.run_count[slot] = count;
                }
                blk_next += count;
                strip_total += count;
                if (c == NC - 1) {
                    if (lane == 0 && strip_total != 0u) atomicAdd(&p.item_count[cur], strip_total);
                    strip_total = 0u;
                }
                __syncwarp();
                // the last emit warp to get here wipes the chunk's cells and gives plane and list back
                uint32_t last = 0u;
                if (lane == 0) {
                    __threadfence_block();
                    last = atomicAdd(&e_done[pi], 1u) == (uint32_t)(kEmitWarps - 1) ? 1u : 0u;
                }
                last = __shfl_sync(0xffffffffu, last, 0);
                if (last) {
                    __threadfence_block();
                    cell_t *wplane = planes + pi * (SR * kPlaneW);
                    if (kn <= (uint32_t)kKlistCap) {
                        for (uint32_t i = (uint32_t)lane; i < kn; i += 32u) {
                            const uint32_t ent = klist[i];
                            wplane[(ent >> 8) * kPlaneW + (ent & 0xffu) - kPlaneLead] = (cell_t)0;
                        }
                    } else {
                        uint4 *pz = reinterpret_cast<uint4 *>(wplane);
                        for (int i = lane; i < L.plane_bytes / 16; i += 32) pz[i] = make_uint4(0u, 0u, 0u, 0u);
                    }
                    __syncwarp();
                    if (lane == 0) {
                        kcount[pi] = 0u;
                        e_done[pi] = 0u;
                        __threadfence_block();
                        mbar_arrive(&p_free[pi]);
                    }
                }
                FDF_CLK(11)
                next_chunk();
            }
        }
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

__global__ void __launch_bounds__(kGatherThreads) fdf_gather_kernel(const DetectParams p, uint32_t n_items) {
    extern __shared__ __align__(16) uint32_t gsm[];
    __shared__ uint32_t warp_sums[kGatherThreads / 32];
    __shared__ unsigned long long s_run_base[kGatherMaxRuns];
    __shared__ uint32_t s_run_count[kGatherMaxRuns];
    __shared__ StripRecord s_rec;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int mode = (int)p.mode, sr = (int)p.sr;
    const int WW = (int)p.words_per_row, NR = (int)p.chunks_per_strip * kRunsPerChunk;  // run records per strip
    const int nwords = out_rows(mode, sr) * WW;
    const int nsum = (nwords + 31) / 32;  // level-2 words
    uint32_t *bits = gsm, *summary = gsm + nsum * 32;
    for (int i = tid; i < nsum * 33; i += kGatherThreads) gsm[i] = 0u;

    // records of a strip: thread r < NR holds run r, thread NR the strip's total and destination
    unsigned long long r_base = 0ull;
    uint32_t r_count = 0u;
    auto fetch = [&](uint32_t item) {
        r_base = 0ull;
        r_count = 0u;
        if (item >= n_items) return;
        if (tid < NR) {
            const size_t slot = (size_t)item * NR + tid;
            r_count = p.run_count[slot];
            r_base = r_count != 0u ? p.run_base[slot] : 0ull;
        } else if (tid == NR) {
            r_count = p.item_count[item];
            r_base = p.item_dst[item];
        }
    };
    fetch(blockIdx.x);
    // sharded batch: where this rank's points start in the batch result, and (first CTA) the global CSR offsets
    unsigned long long shard_base = 0ull;
    if (p.all_offsets != nullptr) {
        for (uint32_t r = 0; r < p.shard_rank; r++) {
            const uint32_t fr = shard_lo(p.total_frames, r + 1u, p.shard_ranks) - shard_lo(p.total_frames, r, p.shard_ranks);
            shard_base += p.all_offsets[(size_t)r * p.shard_block + fr];
        }
        if (blockIdx.x == 0) {
            unsigned long long rb = 0ull;
            for (uint32_t r = 0; r < p.shard_ranks; r++) {
                const uint32_t lo = shard_lo(p.total_frames, r, p.shard_ranks);
                const uint32_t fr = shard_lo(p.total_frames, r + 1u, p.shard_ranks) - lo;
                const unsigned long long *blk = p.all_offsets + (size_t)r * p.shard_block;
                for (uint32_t f = (uint32_t)tid; f < fr; f += kGatherThreads) p.global_offsets[lo + f] = rb + blk[f];
                rb += blk[fr];
            }
            if (tid == 0) p.global_offsets[p.total_frames] = rb;
        }
    }
    // (row, column) of the first level-1 word of each round of this thread: fixed for the whole kernel
    const int row_first = (32 * tid) / WW, col_first = 32 * tid - row_first * WW;
    const int row_step = (32 * kGatherThreads) / WW, col_step = 32 * kGatherThreads - row_step * WW;
    for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        __syncthreads();  // the previous strip is finished with the records (and, the first time, the bitmap is zero)
        if (tid < NR) {
            s_run_base[tid] = r_base;
            s_run_count[tid] = r_count;
        } else if (tid == NR) {
            s_rec.dst = r_base;
            s_rec.total = r_count;
        }
        __syncthreads();
        fetch(item + gridDim.x);  // in flight while this strip is processed
        if (s_rec.total == 0u) continue;  // (block-uniform)
        const uint32_t strip = item % p.strips_per_frame;
        const uint32_t y0 = (uint32_t)(first_out_row(mode) + (int)strip * out_rows(mode, sr));
        // one warp per run: its entries -> bits
        for (int c = warp; c < NR; c += kGatherThreads / 32) {
            const unsigned long long base = s_run_base[c];
            const uint32_t cnt = s_run_count[c];
            for (uint32_t i = (uint32_t)lane; i < cnt; i += 32u) {
                if (base + i >= p.staging_cap) {  // (the emit warp that dropped these entries has raised the flag too)
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
        const unsigned long long o = s_rec.dst + shard_base;
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

size_t detect_smem_bytes(int mode, int sr) { return (size_t)layout_sizes(mode, sr).total; }

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
    if (grid > items) grid = items;
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

cudaError_t launch_scan(const DetectParams &p, cudaStream_t stream) {
    const unsigned long long items = (unsigned long long)p.n_frames * p.strips_per_frame;
    if (items == 0 || items > 0x7fffffffull) return cudaErrorInvalidValue;
    const unsigned tiles = (unsigned)((items + kScanTile - 1) / kScanTile);
    fdf_scan_kernel<<<tiles, kScanThreads, 0, stream>>>(p, (uint32_t)items);
    return cudaGetLastError();
}

cudaError_t launch_gather(const DetectParams &p, cudaStream_t stream, DeviceInfo &info) {
    const unsigned long long items = (unsigned long long)p.n_frames * p.strips_per_frame;
    // (a rank without frames still launches one CTA in a sharded call: it writes the batch's global offsets)
    if ((items == 0 && p.all_offsets == nullptr) || items > 0x7fffffffull) return cudaErrorInvalidValue;
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
    if (grid == 0) grid = 1;
    fdf_gather_kernel<<<(unsigned)grid, kGatherThreads, smem, stream>>>(p, (uint32_t)items);
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
