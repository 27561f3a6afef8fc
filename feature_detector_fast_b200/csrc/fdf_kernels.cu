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
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %4;\n"
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
#ifndef FDF_ABLATE
#define FDF_ABLATE 0  // (timing experiments only, results are wrong: skip phase B = 1, the NMS pass = 2, phase A = 4,
                      //  B arithmetic = 8, B ring loads = 16, stage 2 = 32, the candidate push = 64)
#endif
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

template <int MODE, int SR>
struct Layout {
    static constexpr int TR = tile_rows(SR);
    static constexpr int tile_bytes = TR * kTileW;  // one TMA box
    static constexpr int plane_off = kTileStages * tile_bytes;
    static constexpr int plane_bytes = SR * kPlaneW * 2;  // u16: tag << 12 | score (Off mode: score 1)
    static constexpr int queue_off = plane_off + plane_bytes;
    static constexpr int queue_bytes = kQueueBufs * kQueueCap * 2;      // candidate queues, per chunk
    static constexpr int klist_off = queue_off + queue_bytes;           // keypoint list of the chunk being tested
    static constexpr int klist_bytes = kQueueCap * 2;
    static constexpr int wq_off = klist_off + klist_bytes;              // filter warps' stage-1 -> stage-2 queues
    static constexpr int wq_bytes = 4 * kWarpQueueCap * 2;  // (a filter warp looks at 32 * SR / (2 kFilterWarps) <= 1024 / kFilterWarps groups)
    static constexpr int vtab_off = wq_off + wq_bytes;                  // validity tables: first / middle / last chunk
    static constexpr int vtab_bytes = 3 * kVtabWords * 4;
    static constexpr int misc_off = vtab_off + vtab_bytes;
    static constexpr int misc_bytes = 192;
    static constexpr int total = misc_off + misc_bytes;
    static_assert(tile_bytes % 128 == 0, "TMA destination must stay 128-byte aligned");
    static_assert(plane_bytes % 16 == 0 && plane_off % 16 == 0, "the plane is cleared with 128-bit stores");
    static_assert(SR % (2 * kFilterWarps) == 0 && SR <= 64, "filter warps take row pairs; queue entries hold 6 row bits");
    static_assert(kFallbackWarps * kWarpQueueCap * 2 <= kQueueCap * 2 && kFallbackWarps <= kTestWarps,
                  "the dense fallback borrows the keypoint list buffer for its warp queues");
    static_assert(SR % (2 * kFallbackWarps) == 0, "fallback filter warps take row pairs");
    static_assert(vtab_off % 16 == 0, "the validity tables are read with 128-bit loads");
    static_assert(kGroupRows * kTileW <= kQueueCap, "a row group must always fit the candidate queue");
    static_assert(SR % kGroupRows == 0, "the dense fallback walks whole row groups");
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

// Optional phase clocks (tools/phase_clocks.py builds the library with -DFDF_PHASE_CLOCKS): cycles per phase, summed over
// the warps of a group (lane 0 of each) and over all CTAs.
#ifdef FDF_PHASE_CLOCKS
__device__ unsigned long long g_phase_clocks[16 * 16];  // [warp][slot]
#define FDF_CLK_BEGIN                 \
    long long clk_prev = clock64();   \
    long long clk_acc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#define FDF_CLK(slot)                              \
    {                                              \
        const long long clk_now = clock64();       \
        clk_acc[slot] += clk_now - clk_prev;       \
        clk_prev = clk_now;                        \
    }
#define FDF_CLK_END                                                                               \
    if (lane == 0)                                                                                \
        for (int i = 0; i < 10; i++)                                                              \
            if (clk_acc[i]) atomicAdd(&g_phase_clocks[warp * 16 + i], (unsigned long long)clk_acc[i]);
#elif defined(FDF_TRACE)
// Timeline trace (tools/trace_timeline.py builds the library with -DFDF_TRACE): the CTAs resident on SM 0 write
// clock64() at every phase boundary of their first kTraceChunks chunks, per warp (lane 0), into a global table.
constexpr int kTraceCtas = 4, kTraceChunks = 200, kTraceSlots = 12;
__device__ long long g_trace[kTraceCtas][16][kTraceChunks][kTraceSlots];
__device__ unsigned int g_trace_n;
#define FDF_CLK_BEGIN
#define FDF_CLK(slot)                                                                                           \
    if (trace_cta >= 0 && lane == 0 && gc < (uint32_t)kTraceChunks) g_trace[trace_cta][warp][gc][slot] = clock64();
#define FDF_CLK_END
#else
#define FDF_CLK_BEGIN
#define FDF_CLK(slot)
#define FDF_CLK_END
#endif
#ifdef FDF_TRACE
#define FDF_TRACE_ROW \
    ((trace_cta >= 0 && lane == 0 && gc < (uint32_t)kTraceChunks) ? &g_trace[trace_cta][warp][gc][0] : nullptr)
#else
#define FDF_TRACE_ROW nullptr
#endif

// ---- the detection kernel ----------------------------------------------------------------------
// Two groups of warps per CTA, coupled only through mbarriers (no CTA-wide barrier inside the chunk loop):
//   filter warps (0 .. kFilterWarps-1): wait for chunk k's tile, run phase A into candidate queue k % 3,
//       arrive on q_full[k % 3], go on to chunk k + 1;
//   test warps (the others): wait on q_full[k % 3], run phase B, barrier among themselves, then one thread
//       requests the tile of chunk k + 2 (the tile buffer of chunk k is free now), all run the NMS pass over
//       the chunk's keypoint list, barrier, and the last warp copies the survivors to the staging buffer
//       while the others already wait for chunk k + 1.
// A landed tile k + 2 implies that chunk k - 1 is completely done, hence that queue (k + 2) % 3 is free again.
// Timing experiment (-DFDF_QFULL_BAR): queue hand-off filter -> test warps on a hardware named barrier (ids 4 ..): the
// filter warps only ARRIVE, the test warps SYNC, i.e. they block in hardware instead of polling an mbarrier (the q_full
// polls are 15 % of all executed instructions).  Measured 1.9 % SLOWER (1.035 vs 1.015 ms per 256 frames): the polls
// use issue slots nobody else wants, and a polled wait is left sooner than a named barrier.  Default: mbarrier.
__device__ __forceinline__ void qfull_arrive(uint32_t qb) {  // (immediate barrier ids: a register id reserves all 16)
    if (qb == 0u) asm volatile("bar.arrive 4, %0;" ::"n"(kThreads) : "memory");
    else if (qb == 1u) asm volatile("bar.arrive 5, %0;" ::"n"(kThreads) : "memory");
    else if (qb == 2u) asm volatile("bar.arrive 6, %0;" ::"n"(kThreads) : "memory");
    else asm volatile("bar.arrive 7, %0;" ::"n"(kThreads) : "memory");
}
__device__ __forceinline__ void qfull_sync(uint32_t qb) {
    if (qb == 0u) asm volatile("bar.sync 4, %0;" ::"n"(kThreads) : "memory");
    else if (qb == 1u) asm volatile("bar.sync 5, %0;" ::"n"(kThreads) : "memory");
    else if (qb == 2u) asm volatile("bar.sync 6, %0;" ::"n"(kThreads) : "memory");
    else asm volatile("bar.sync 7, %0;" ::"n"(kThreads) : "memory");
}

__device__ __forceinline__ void bar_test_group() {
    asm volatile("bar.sync 1, %0;" ::"n"(kTestThreads) : "memory");
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
#ifndef FDF_SMALL_SR_CTAS
#define FDF_SMALL_SR_CTAS 4
#endif
__global__ void __launch_bounds__(kThreads, SR >= 48 ? 3 : FDF_SMALL_SR_CTAS)
fdf_detect_kernel(const __grid_constant__ CUtensorMap tmap, const DetectParams p) {
    using L = Layout<MODE, SR>;
    constexpr int HS = (MODE == NMS_OFF) ? 0 : 1;  // score halo (rows and columns) needed by the 3x3 NMS
    constexpr int OUT_R = out_rows(MODE, SR);

    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *tiles = smem;
    uint16_t *plane = reinterpret_cast<uint16_t *>(smem + L::plane_off);
    uint16_t *queues = reinterpret_cast<uint16_t *>(smem + L::queue_off);              // [kQueueBufs][kQueueCap]
    uint16_t *klist = reinterpret_cast<uint16_t *>(smem + L::klist_off);               // [kQueueCap]
    uint16_t *wqs = reinterpret_cast<uint16_t *>(smem + L::wq_off);                    // [kFilterWarps][kWarpQueueCap]
    uint32_t *vtabs = reinterpret_cast<uint32_t *>(smem + L::vtab_off);                 // [3][kVtabWords]
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + L::misc_off);             // [kTileStages] tile landed
    uint64_t *q_full = reinterpret_cast<uint64_t *>(smem + L::misc_off + 32);          // [kQueueBufs] candidate queue complete
    uint32_t *qcount = reinterpret_cast<uint32_t *>(smem + L::misc_off + 64);          // [kQueueBufs] queue fill
    uint32_t *s_ticket = reinterpret_cast<uint32_t *>(smem + L::misc_off + 80);        // [2] strip tickets (strip parity)
    uint32_t *scount = reinterpret_cast<uint32_t *>(smem + L::misc_off + 88);          // keypoints staged by the chunk so far
    uint32_t *s_total = reinterpret_cast<uint32_t *>(smem + L::misc_off + 92);         // keypoints of the strip so far
    unsigned long long *s_base = reinterpret_cast<unsigned long long *>(smem + L::misc_off + 96);
    unsigned long long *s_block = reinterpret_cast<unsigned long long *>(smem + L::misc_off + 104);  // [2] staging block
    volatile uint32_t *s_abort = reinterpret_cast<volatile uint32_t *>(smem + L::misc_off + 120);    // a wait timed out
    uint32_t *kcount = reinterpret_cast<uint32_t *>(smem + L::misc_off + 124);         // [2] keypoint list fill (chunk parity)
    [[maybe_unused]] volatile long long *clk_req =
        reinterpret_cast<volatile long long *>(smem + L::misc_off + 136);              // [kTileStages] (phase clocks only)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int W = (int)p.w, H = (int)p.h;
    const int NC = (int)p.chunks_per_strip;
    const uint32_t total_items = p.n_frames * p.strips_per_frame;
    const bool is_filter = warp < kFilterWarps;
    const int ttid = tid - kFilterWarps * 32;  // thread index inside the test group
#ifndef FDF_T0_WARP
#define FDF_T0_WARP (kTestWarps - 1)
#endif
    // the thread that draws tickets, requests tiles and keeps the run records: lane 0 of the LAST test warp, which gets
    // the smallest share of every candidate / keypoint list -- its serial work between the group's barriers then
    // overlaps the other warps' list work instead of extending the critical path
    const bool t0 = ttid == 32 * FDF_T0_WARP;

    uint32_t cur = 0u, nxt = 0xffffffffu;
    bool have_nxt = false;
    const int ahead = NC >= kTileStages ? kTileStages : NC;  // tiles requested this many chunks ahead (never beyond the next strip)

    // request the tile of chunk c of the current strip (c < NC) or of chunk c - NC of the next strip; `it` is the
    // current strip's sequence number in this CTA.  When the work is exhausted the barrier is completed without
    // a tile, so that the filter warps wake up and see the end ticket.
    auto request_tile = [&](int c, uint32_t stream_index, uint32_t it) {
        uint32_t item = cur;
        if (c >= NC) {
            if (c - NC >= NC) return;
            if (!have_nxt) {
                nxt = atomicAdd(p.ticket, 1u);
                have_nxt = true;
                s_ticket[(it + 1u) & 1u] = nxt;
            }
            item = nxt;
            c -= NC;
        }
        const uint32_t stage = stream_index % (uint32_t)kTileStages;
        if (item >= total_items) {
            mbar_arrive(&full_bar[stage]);
            return;
        }
        const uint32_t frame = item / p.strips_per_frame;
        const uint32_t strip = item - frame * p.strips_per_frame;
        const int ty0 = first_out_row(MODE) + (int)strip * OUT_R - HS - 3;  // image row of tile row 0
#ifdef FDF_PHASE_CLOCKS
        clk_req[stage] = clock64();
#endif
        mbar_expect_tx(&full_bar[stage], (uint32_t)L::tile_bytes);
        tma_load_3d(tiles + stage * L::tile_bytes, &tmap, c * kChunkW - kTileLead, ty0, (int)frame, &full_bar[stage]);
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
        kcount[0] = kcount[1] = 0u;
        s_block[0] = s_block[1] = 0ull;
        *s_base = open_run(s_block, p);
        cur = atomicAdd(p.ticket, 1u);
        s_ticket[0] = cur;
    }
    if (tid < 3 * kVtabWords) vtabs[tid] = valid_word<MODE>(W, vtab_chunk(tid / kVtabWords, NC), tid % kVtabWords);
    auto clear_plane = [&](int i0, int n) {  // by n threads, i0 = index of this one
        uint4 *pz = reinterpret_cast<uint4 *>(plane);
        for (int i = i0; i < L::plane_bytes / 16; i += n) pz[i] = make_uint4(0u, 0u, 0u, 0u);
    };
    clear_plane(tid, kThreads);
#ifdef FDF_TRACE
    volatile int &s_trace_cta = *reinterpret_cast<volatile int *>(smem + L::misc_off + 176);
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
        for (int c = 0; c < ahead; c++) request_tile(c, (uint32_t)c, 0u);  // (ahead <= NC: all of the first strip)

    const int t = (int)p.threshold, n = (int)p.count;
    const uint32_t kbias = filter_kbias(p.threshold);
    uint32_t gc = 0;   // chunks processed by this CTA so far
    uint32_t qb = 0, qpar = 0;  // tile stage = queue buffer = gc % kTileStages, and the parity of their mbarrier phases

    if (is_filter) {
        // ================================ filter warps =================================================
        uint16_t *wq = wqs + warp * (4 * kWarpQueueCap / kFilterWarps);
        FDF_CLK_BEGIN
        for (uint32_t it = 0; cur < total_items; it++) {
            const uint32_t frame = cur / p.strips_per_frame;
            const uint32_t strip = cur - frame * p.strips_per_frame;
            for (int c = 0; c < NC; c++, gc++) {
                const uint32_t stage = qb;
                const ChunkGeo g = make_geo<MODE>(W, H, (int)strip, c, SR);
#ifdef FDF_PHASE_CLOCKS
                uint32_t landed;
                asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                             : "=r"(landed) : "r"(smem_u32(&full_bar[stage])), "r"(qpar) : "memory");
                const bool had_to_wait = landed == 0u;
#endif
                mbar_wait(&full_bar[stage], qpar, p.flags, s_abort);  // (also: queue qb is free again)
#ifdef FDF_PHASE_CLOCKS
                if (had_to_wait) {
                    clk_acc[8] += clock64() - clk_req[stage];
                    clk_acc[9] += 1;
                }
#endif
                FDF_CLK(0)
#if !(FDF_ABLATE & 4)
                phase_a_warp<MODE, SR, kFilterWarps>(warp, lane, tiles + stage * L::tile_bytes, wq,
                                                     vtabs + vtab_variant(c, NC) * kVtabWords, vtab_variant(c, NC),
                                                     queues + qb * kQueueCap,
                                                     &qcount[qb], g, kbias, 0, SR, FDF_TRACE_ROW);
#endif
                __syncwarp();
#ifndef FDF_QFULL_BAR
                if (lane == 0) mbar_arrive(&q_full[qb]);
#else
                qfull_arrive(qb);
#endif
                FDF_CLK(1)
                if (++qb == (uint32_t)kQueueBufs) {
                    qb = 0;
                    qpar ^= 1u;
                }
            }
            // the next strip's ticket is published before its first tile is requested (or the end is signalled)
            mbar_wait(&full_bar[qb], qpar, p.flags, s_abort);
            FDF_CLK(2)
            cur = s_ticket[(it + 1u) & 1u];
        }
        FDF_CLK_END
        return;
    }

    // ==================================== test warps ===================================================
    uint32_t tag = 1u;  // (gc % kTagPeriod) + 1
    const int twarp = warp - kFilterWarps;
    // t0, between the group's barriers: the chunk's run record and the strip's running total
    auto close_run = [&](uint32_t slot, uint32_t count, bool last) {
        if (count != 0u) {
            p.run_base[slot] = *s_base;
            p.run_count[slot] = count;
        }
        p.run_n[slot] = count != 0u ? 1u : 0u;
        const uint32_t tot = *s_total + count;
        if (last) p.item_count[cur] = tot;
        *s_total = last ? 0u : tot;
    };
    FDF_CLK_BEGIN
    for (uint32_t it = 0; cur < total_items; it++) {
        const uint32_t frame = cur / p.strips_per_frame;
        const uint32_t strip = cur - frame * p.strips_per_frame;
        for (int c = 0; c < NC; c++, gc++) {
            const uint32_t stage = qb;
            const uint8_t *tile = tiles + stage * L::tile_bytes;
            uint16_t *queue = queues + qb * kQueueCap;
            const ChunkGeo g = make_geo<MODE>(W, H, (int)strip, c, SR);
            const uint32_t slot = cur * (uint32_t)NC + (uint32_t)c;
            if (tag == 1u && gc != 0u) {  // every 15 chunks: restart the tags on a cleared plane
                clear_plane(ttid, kTestThreads);
                bar_test_group();
            }
#ifndef FDF_QFULL_BAR
            mbar_wait(&q_full[qb], qpar, p.flags, s_abort);
#else
            qfull_sync(qb);
#endif
            mbar_wait(&full_bar[stage], qpar, p.flags, s_abort);  // (completed long ago: makes the tile visible here too)
            FDF_CLK(4)
            const uint32_t qn = qcount[qb];
            if (qn <= (uint32_t)kQueueCap) {
                // Timing experiment (-DFDF_EARLY_TILE_REQUEST, measured 4.5 % SLOWER: 1.069 vs 1.022 ms per 256 frames):
                // the tile and the candidate queue of this chunk are free as soon as every test thread has loaded
                // the ring bytes of its last candidate, so the other warps only ARRIVE on named barrier 2 there and
                // the housekeeping warp waits on it and requests the tile of chunk k + 2 before the arithmetic of
                // the last step.  Default: the request follows the group barrier after phase B.
                auto tile_done = [&]() {
#ifdef FDF_EARLY_TILE_REQUEST
                    if (twarp == FDF_T0_WARP) {
                        asm volatile("bar.sync 2, %0;" ::"n"(kTestThreads) : "memory");
                        if (t0) {
                            qcount[qb] = 0u;
                            request_tile(c + ahead, gc + (uint32_t)ahead, it);
                        }
                    } else {
                        asm volatile("bar.arrive 2, %0;" ::"n"(kTestThreads) : "memory");
                    }
#endif
                };
#if !(FDF_ABLATE & 1)
                phase_b<MODE, SR, kTestUnroll>(ttid, lane, kTestThreads, qn, tile, queue, klist, &kcount[gc & 1u], plane, t, n, tag,
                                               tile_done);
#else
                tile_done();
#endif
                FDF_CLK(5)
                bar_test_group();  // every score of this chunk is in the plane and its keypoint list is complete
                FDF_CLK(3)
#if !defined(FDF_EARLY_TILE_REQUEST) && !defined(FDF_LATE_TILE_REQUEST)
                if (t0) {
                    qcount[qb] = 0u;
                    request_tile(c + ahead, gc + (uint32_t)ahead, it);
                }
#endif
                FDF_CLK(6)
#if !(FDF_ABLATE & 2)
                emit_list<MODE, SR>(ttid, kTestThreads, kcount[gc & 1u], klist, plane, scount, *s_base, p.staging_cap,
                                    p.staging, g);
#endif
                FDF_CLK(7)
                bar_test_group();  // the run is complete; the keypoint list is free
#ifdef FDF_LATE_TILE_REQUEST  // timing experiment: the filter group gets its next tile only after the NMS pass
                if (t0) {
                    qcount[qb] = 0u;
                    request_tile(c + ahead, gc + (uint32_t)ahead, it);
                }
#endif
                if (t0) {
                    const uint32_t count = *scount;
                    *scount = 0u;
                    kcount[gc & 1u] = 0u;  // (next used two chunks from now)
                    close_run(slot, count, c == NC - 1);
                    s_block[0] = *s_base + count;  // give the unused tail back
                    *s_base = open_run(s_block, p);
                }
                FDF_CLK(8)
            } else {
                // very dense content: redo the chunk kGroupRows rows at a time (the test warps filter for themselves;
                // their warp queues live in the keypoint list buffer, which this path does not use), then
                // emit from the score plane: count, reserve, write
                uint16_t *wq = klist + twarp * kWarpQueueCap;
                const uint32_t *vtab = vtabs + vtab_variant(c, NC) * kVtabWords;
                for (int lo = 0; lo < SR; lo += kGroupRows) {
                    bar_test_group();
                    if (t0) qcount[qb] = 0u;
                    bar_test_group();
                    if (twarp < kFallbackWarps)
                        phase_a_warp<MODE, SR, kFallbackWarps>(twarp, lane, tile, wq, vtab, vtab_variant(c, NC), queue,
                                                               &qcount[qb], g, kbias,
                                                               lo, lo + kGroupRows);
                    bar_test_group();
                    phase_b<MODE, SR>(ttid, lane, kTestThreads, qcount[qb], tile, queue, nullptr, nullptr, plane, t, n, tag);
                }
                bar_test_group();  // every score of this chunk is in the plane; tile[stage] is free again
                if (t0) {
                    qcount[qb] = 0u;
                    request_tile(c + ahead, gc + (uint32_t)ahead, it);
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
                if (kn != 0u) nms_dense<MODE, SR>(ttid, kTestThreads, 1, plane, scount, *s_base, p.staging_cap, p.staging, g, tag);
                bar_test_group();
                if (t0) {
                    *scount = 0u;
                    close_run(slot, kn, c == NC - 1);
                    *s_base = open_run(s_block, p);
                }
            }
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
    FDF_CLK_END
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

    // records of a strip: thread c < NC holds chunk c's run, thread NC the strip's total and destination
    unsigned long long r_base = 0ull;
    uint32_t r_count = 0u;
    auto fetch = [&](uint32_t item) {
        r_base = 0ull;
        r_count = 0u;
        if (item >= n_items) return;
        if (tid < NC) {
            const size_t slot = (size_t)item * NC + tid;
            r_count = p.run_n[slot] != 0u ? p.run_count[slot] : 0u;
            r_base = r_count != 0u ? p.run_base[slot] : 0ull;
        } else if (tid == NC) {
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
                if (base + i >= p.staging_cap) break;
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

#ifdef FDF_PHASE_CLOCKS
}  // namespace
cudaError_t read_phase_clocks(unsigned long long out[256]) {
    cudaError_t e = cudaMemcpyFromSymbol(out, g_phase_clocks, 256 * sizeof(unsigned long long));
    if (e != cudaSuccess) return e;
    unsigned long long zero[256] = {0};
    return cudaMemcpyToSymbol(g_phase_clocks, zero, sizeof(zero));
}
namespace {
#endif

template <int MODE, int SR>
cudaError_t launch_t(const CUtensorMap &tmap, const DetectParams &p, cudaStream_t stream) {
    auto kern = fdf_detect_kernel<MODE, SR>;
    const size_t smem = detect_smem_bytes(MODE, SR);
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
    if (const char *lim = getenv("FDF_CTAS_PER_SM")) {  // tuning knob for experiments
        const int v = atoi(lim);
        if (v >= 1 && v < per_sm) per_sm = v;
    }
    unsigned long long grid = (unsigned long long)sms * (unsigned)per_sm;
    if (grid > items) grid = items;
    kern<<<(unsigned)grid, kThreads, smem, stream>>>(tmap, p);
    return cudaGetLastError();
}

}  // namespace

size_t detect_smem_bytes(int mode, int sr) {
    const size_t tile = (size_t)tile_rows(sr) * kTileW;
    (void)mode;
    const size_t plane = (size_t)sr * kPlaneW * 2;
    const size_t queues = (size_t)(kQueueBufs + 1) * kQueueCap * 2, wq = (size_t)4 * kWarpQueueCap * 2;
    return (size_t)kTileStages * tile + plane + queues + wq + (size_t)3 * kVtabWords * 4 + 192;
}

size_t gather_smem_bytes(int mode, int sr, uint32_t words_per_row) {
    const size_t nsum = ((size_t)out_rows(mode, sr) * words_per_row + 31) / 32;  // level-2 words
    return nsum * 33 * 4;
}
cudaError_t launch_detect(int mode, int sr, const CUtensorMap &tmap, const DetectParams &p, cudaStream_t stream) {
#define FDF_CASE(M, S) \
    if (mode == M && sr == S) return launch_t<M, S>(tmap, p, stream);
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

cudaError_t launch_gather(const DetectParams &p, cudaStream_t stream) {
    const unsigned long long items = (unsigned long long)p.n_frames * p.strips_per_frame;
    // (a rank without frames still launches one CTA in a sharded call: it writes the batch's global offsets)
    if ((items == 0 && p.all_offsets == nullptr) || items > 0x7fffffffull) return cudaErrorInvalidValue;
    const size_t smem = gather_smem_bytes((int)p.mode, (int)p.sr, p.words_per_row);
    cudaError_t e = cudaFuncSetAttribute(fdf_gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int dev = 0, sms = 0, per_sm = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fdf_gather_kernel, kGatherThreads, smem)) != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorInvalidConfiguration;
    unsigned long long grid = (unsigned long long)sms * (unsigned)per_sm;
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
