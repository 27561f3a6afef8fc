// fdf_strip.cuh -- the per-thread / per-warp bodies of the detection kernel's phases.
//
// They are `__host__ __device__` and take the thread (or warp and lane) index as an argument:
// fdf_kernels.cu calls them with threadIdx.x between its barriers, and tests/host/strip_emulator.cpp
// runs the very same code thread by thread on the CPU (TMA replaced by a zero-filled copy, the warp
// ballot replaced by a loop over the 32 lanes) so that the tiling, halo, validity and NMS-row rules
// are checked against the oracle without a GPU.
//
// Geometry (see fdf_kernels.cuh): a strip has SR scored rows; tile row 0 is image row
// ys0 - 3 where ys0 is the image row of scored row 0; tile column 0 is image column xt0,
// where x0 = xt0 + 12 is the chunk's first output column; xt0 = 240*chunk - 16 is a multiple of 16
// because TMA needs the box's innermost start coordinate 16-byte aligned.  A chunk emits columns
// [x0, x1): 240 of them, except that the row's last chunk runs to the image's last centre column.
#pragma once
#include "fdf_core.cuh"
#include "fdf_kernels.cuh"

namespace fdf {

FDF_HD uint32_t atomic_add_u32(uint32_t *p, uint32_t v) {
#if defined(__CUDA_ARCH__)
    return atomicAdd(p, v);
#else
    uint32_t old = *p;
    *p = old + v;
    return old;
#endif
}

FDF_HD int lowest_set_bit(uint32_t m) {  // m != 0
#if defined(__CUDA_ARCH__)
    return __ffs(m) - 1;
#else
    return __builtin_ctz(m);
#endif
}

struct ChunkGeo {
    int w, h;   // image size
    int ys0;    // image row of scored row 0
    int y0;     // first image row this strip emits
    int x0;     // first image column this chunk emits
    int x1;     // one past the last image column this chunk emits
    int xt0;    // image column of tile column 0 (a multiple of 16)
};

template <int MODE>
FDF_HD ChunkGeo make_geo(int w, int h, int strip, int chunk, int sr) {
    ChunkGeo g;
    g.w = w;
    g.h = h;
    g.y0 = first_out_row(MODE) + strip * out_rows(MODE, sr);
    g.ys0 = g.y0 - (MODE == NMS_OFF ? 0 : 1);
    g.xt0 = chunk * kChunkW - kTileLead;
    g.x0 = g.xt0 + kLeftHalo;
    g.x1 = (chunk == chunks_per_row(w) - 1) ? w - 3 : g.x0 + kChunkW;
    return g;
}

// ---- which tile columns may hold a centre (fast_simd.rs:369-371, 559-562) ------------------------------
// Scored columns of a chunk: its own columns plus the one-column score halo the 3x3 NMS needs, inside the
// image's centre range [3, w-3).  The table has one word per 16-pixel group: the group's candidate mask
// (candidate_mask16's bit layout) with every valid centre set -- 16 words, read with one conflict-free LDS by stage 2;
// it depends on the chunk only through "first / middle / last chunk of the row", so the kernel builds the three
// variants once (kVtabWords words each).
constexpr int kVtabWords = kTileW / 16;

// scored tile columns [lo, hi) of a chunk
struct ColRange {
    int lo, hi;
};
template <int MODE>
FDF_HD ColRange scored_cols(int w, int chunk) {
    constexpr int HS = (MODE == NMS_OFF) ? 0 : 1;
    const int xt0 = chunk * kChunkW - kTileLead, x0 = xt0 + kLeftHalo;
    const int x1 = (chunk == chunks_per_row(w) - 1) ? w - 3 : x0 + kChunkW;
    ColRange r;
    r.lo = max(3, x0 - HS) - xt0;
    r.hi = min(w - 3, x1 + HS) - xt0;
    return r;
}

template <int MODE>
FDF_HD uint32_t valid_group_mask(int w, int chunk, int q) {
    const ColRange cr = scored_cols<MODE>(w, chunk);
    uint32_t v = 0u;
    for (int px = 0; px < 16; px++) {
        const int j = 16 * q + px;
        if (j >= cr.lo && j < cr.hi) v |= 1u << px_to_mask_bit(px);
    }
    return v;
}

// variant of chunk c: 2 = last chunk of the row (also when it is the only one), 0 = first, 1 = any other
FDF_HD int vtab_variant(int c, int nc) { return c == nc - 1 ? 2 : (c == 0 ? 0 : 1); }
FDF_HD int vtab_chunk(int variant, int nc) { return variant == 2 ? nc - 1 : variant; }

// ---- phase A: two-stage dense filter (replaces fast_simd.rs:368-520) -------------------------------------
// Stage 1 looks at every scored pixel, 16 per lane and row, north/south pair only (~12 % of the pixels of
// natural content pass, ~19 % of the 16-pixel groups).  Groups with a survivor are compacted into the warp's
// segment of the chunk's entry table (ballot + rank, no atomics); after a barrier among the filter warps,
// stage 2 runs the full two-pair filter with one lane per entry -- over the entries of ALL warps, dealt round
// robin to the filter threads, because a horizontal edge puts most of a chunk's entries into one warp's rows --
// and pushes every surviving centre (~2 % of the pixels) to the chunk's candidate queue.
//
// Entry (one byte): row inside the warp's rows << 4 | group.  Candidate: scored row << 9 | group << 5 | mask bit.

// scored rows [lo, hi) of a strip that can hold a centre at all (fast_simd.rs:342: image rows 3 .. h-4)
struct RowRange {
    int lo, hi;
};
FDF_HD RowRange live_rows(const ChunkGeo &g, int sr) {
    RowRange r;
    r.lo = max(0, 3 - g.ys0);
    r.hi = min(sr, g.h - 3 - g.ys0);
    return r;
}

// Stage 1 for one lane: the 16-pixel group q of BH consecutive scored rows rr0 .. rr0+BH-1.  Returns a mask with
// bit i set iff row rr0 + i has a centre whose north or south ring pixel differs from it by more than t.
// The rows are walked top to bottom with everything kept in registers: every tile row is loaded once (LDS.128)
// and the vertical difference D(y) = |p(y) - p(y-3)| is computed once and used twice (as the north difference of
// centre y and the south difference of centre y - 3).
template <int BH>
FDF_HD uint32_t stage1_band(const uint8_t *tile, int rr0, int q, uint32_t kbias) {
    const uint8_t *p = tile + rr0 * kTileW + q * 16;  // tile row rr0 is the north ring row of scored row rr0
    Px16 row[BH + 6], d[BH + 3];
    uint32_t v = 0u;
#pragma unroll
    for (int i = 0; i < BH + 6; i++) {
        row[i] = load16(p + i * kTileW);
        if (i >= 3) {
#pragma unroll
            for (int k = 0; k < 4; k++) d[i - 3].w[k] = absdiff4(row[i].w[k], row[i - 3].w[k]);
        }
        if (i >= 6) {  // centre = tile row i - 3: north difference d[i - 6], south difference d[i - 3]
            // "some centre of the group has a north or south difference > t" is only needed per group, and stage 2
            // repeats the test per pixel: so OR the eight difference words first and threshold once.  A byte of the OR
            // is >= every byte that went into it, so no exceeding difference is lost; the OR of two differences <= t
            // can exceed t, which only sends a few more groups to stage 2 (18.8 % instead of 18.5 % at t = 20).
            uint32_t x = d[i - 6].w[0] | d[i - 3].w[0];
#pragma unroll
            for (int k = 1; k < 4; k++) x |= d[i - 6].w[k] | d[i - 3].w[k];
            if ((exceeds4(x, kbias) & 0x80808080u) != 0u) v |= 1u << (i - 6);
        }
    }
    return v;
}

// bit i set iff scored row rr0 + i may hold a centre at all (fast_simd.rs:342: image rows 3 .. h-4)
FDF_HD uint32_t live_mask(const ChunkGeo &g, int rr0, int bh, int sr) {
    const RowRange live = live_rows(g, sr);
    uint32_t m = 0u;
    for (int i = 0; i < bh; i++)
        if (rr0 + i >= live.lo && rr0 + i < live.hi) m |= 1u << i;
    return m;
}

// One filter warp's stage 1; NW warps share a chunk.  Lane l of warp v handles 16-pixel group q = l & 15 of the
// BH = SR / (2 NW) scored rows starting at (2 v + (l >> 4)) * BH.  The warp's entries go to `ent` (its segment of the
// chunk's entry table, kWarpQueueCap bytes); returns how many.  On the device the 32 lanes run it together; on the host
// the emulator calls it once per warp (lane = -1) and the lane loop runs sequentially, in the order of the ballot ranks.
template <int MODE, int SR, int NW>
FDF_HD uint32_t phase_a_stage1(int warp, int lane_or_minus1, const uint8_t *tile, uint8_t *ent, const ChunkGeo &g,
                               uint32_t kbias) {
    constexpr int BH = SR / (2 * NW);
    static_assert(2 * BH <= 16, "an entry holds 4 bits of row");
    static_assert(BH * 32 <= kWarpQueueCap, "a warp's segment must hold every group stage 1 looks at");
    uint32_t n = 0u;  // entries (warp-uniform)
#if defined(__CUDA_ARCH__)
    const int lane = lane_or_minus1;
    const int q = lane & 15, half = lane >> 4, rr0 = (2 * warp + half) * BH;
    const uint32_t lt = (1u << lane) - 1u;
    const uint32_t need = stage1_band<BH>(tile, rr0, q, kbias) & live_mask(g, rr0, BH, SR);
    const uint32_t e0 = (uint32_t)(((half * BH) << 4) | q);
#pragma unroll
    for (int i = 0; i < BH; i++) {
        const bool mine = (need >> i) & 1u;
        const uint32_t b = __ballot_sync(0xffffffffu, mine);
        if (mine) {
            FDF_BOUND(n + (uint32_t)__popc(b & lt), kWarpQueueCap);
            ent[n + (uint32_t)__popc(b & lt)] = (uint8_t)(e0 + (uint32_t)(i << 4));
        }
        n += (uint32_t)__popc(b);
    }
#else
    (void)lane_or_minus1;
    uint32_t need[32];
    for (int lane = 0; lane < 32; lane++) {
        const int rr0 = (2 * warp + (lane >> 4)) * BH;
        need[lane] = stage1_band<BH>(tile, rr0, lane & 15, kbias) & live_mask(g, rr0, BH, SR);
    }
    for (int i = 0; i < BH; i++)
        for (int lane = 0; lane < 32; lane++)
            if ((need[lane] >> i) & 1u) {
                FDF_BOUND(n, kWarpQueueCap);
                ent[n++] = (uint8_t)(((((lane >> 4) * BH) + i) << 4) | (lane & 15));
            }
#endif
    return n;
}

// stage 2 for one 16-pixel group (scored row rr, group q): the candidate mask of its 16 centres (bit layout:
// candidate_mask16)
FDF_HD uint32_t stage2_mask(int rr, int q, const uint8_t *tile, const uint32_t *vtab, uint32_t kbias) {
    const uint8_t *rowp = tile + (rr + 3) * kTileW + q * 16;
    // the words left / right of the group: for q = 0 / q = 15 they belong to the neighbouring tile row, which
    // only reaches centres the validity table excludes (tile columns 0..2 and 253..255)
    const uint32_t cl = *reinterpret_cast<const uint32_t *>(rowp - 4);
    const uint32_t cr = *reinterpret_cast<const uint32_t *>(rowp + 16);
    return candidate_mask16(load16(rowp), load16(rowp - 3 * kTileW), load16(rowp + 3 * kTileW), cl, cr, vtab[q], kbias);
}

// one queue entry per set bit of the group's candidate mask, at queue[slot ...]
FDF_HD void push_candidates(int rr, int q, uint32_t m, uint16_t *queue, uint32_t slot) {
    const uint32_t base = (uint32_t)(rr << 9 | q << 5);
    uint16_t *out = queue + slot;
    while (m != 0u) {
        const uint32_t p = (uint32_t)highest_set_bit(m);
        m ^= 1u << p;
        FDF_BOUND(out - queue, kQueueCap);
        *out++ = (uint16_t)(base + p);
    }
}

// Stage 2 for one filter thread (ftid of nthreads): entries ftid, ftid + nthreads, ... of the chunk's entry table, the
// warps' segments taken one after the other (nent[v] entries in segment v).  When the candidate queue is full the
// entries are dropped but still counted: *qcount > kQueueCap tells the test warps to take the dense path.
template <int MODE, int SR, int NW>
FDF_HD void phase_a_stage2(int ftid, int nthreads, const uint8_t *tile, const uint8_t *ents, const uint32_t *nent,
                           const uint32_t *vtab, uint16_t *queue, uint32_t *qcount, uint32_t kbias) {
    constexpr int RW = SR / NW;  // scored rows per filter warp
    uint32_t cnt[NW], total = 0u;
#pragma unroll
    for (int v = 0; v < NW; v++) {
        cnt[v] = nent[v];
        total += cnt[v];
    }
    // entry gidx of the concatenated segments -> (scored row, group)
    auto locate = [&](uint32_t i, int &rr, int &q) {
        int v = 0;
#pragma unroll
        for (int u = 0; u < NW - 1; u++)
            if (v == u && i >= cnt[u]) {
                i -= cnt[u];
                v = u + 1;
            }
        FDF_BOUND(i, kWarpQueueCap);
        const uint32_t e = ents[v * kWarpQueueCap + (int)i];
        rr = v * RW + (int)(e >> 4);
        q = (int)(e & 15u);
        FDF_BOUND(rr, SR);
        FDF_BOUND((rr + 6) * kTileW + q * 16 + 15, tile_rows(SR) * kTileW);  // the group's south ring row is in the tile
    };
    for (uint32_t gidx = (uint32_t)ftid; gidx < total; gidx += (uint32_t)nthreads) {
        int rr, q;
        locate(gidx, rr, q);
        const uint32_t m = stage2_mask(rr, q, tile, vtab, kbias);
        if (m != 0u) {
            const uint32_t c = (uint32_t)popc32(m);
            const uint32_t slot = atomic_add_u32(qcount, c);
            if (slot + c <= (uint32_t)kQueueCap) push_candidates(rr, q, m, queue, slot);
        }
    }
}

// ---- phase B: exact segment test (+ score) per candidate (replaces fast_simd.rs:115-297, 623-749)
// One thread per queue entry: 16 ring bytes + the centre from the tile, one dual word per ring pixel (fdf_core.cuh),
// best window -> keypoint yes / no and the MaxThreshold score in the same ~50 instructions.  Every keypoint writes
// (tag << 12 | score) into the score plane at (scored row, tile column - kPlaneLead) -- in Off mode the score is 1
// and only the dense path reads it -- and is appended to the chunk's keypoint list as scored row << 8 | tile
// column (one ballot per warp step, one shared atomic per warp step that found a keypoint).  klist == nullptr
// (dense path): no list.
// On the device the 32 lanes of a warp call it together (lane >= 0); the host emulator calls it once per thread
// with lane = -1.
struct KeypointTest {
    bool kp;
    uint32_t score;
};

// NFIX: the consecutive count as a compile-time constant (9), or 0 = use n (see best_window)
template <int MODE, int NFIX>
FDF_HD KeypointTest test_pixel(const uint8_t *pc, int t, int n) {
    RingDual ring;
    const uint32_t bias = dual_bias((int)pc[0]);
#pragma unroll
    for (int k = 0; k < 16; k++) ring.w[k] = dual_word((uint32_t)pc[FDF_RING_DY(k) * kTileW + FDF_RING_DX(k)], bias);
    const uint32_t best = best_of_lanes(best_window<NFIX>(ring, n));
    KeypointTest r;
    r.kp = best > (uint32_t)(256 + t);
    r.score = 1u;  // Off mode: the plane only records "keypoint here" (used by the dense path)
    if (MODE == NMS_MAX_THRESHOLD) r.score = best - 256u;                  // (garbage unless kp)
    if (MODE == NMS_SUM_ABSOLUTE) r.score = score_sum_abs_dual(ring, t);  // <= 4080 < 2^12
    return r;
}

template <int MODE, int SR, int NFIX>
FDF_HD void phase_b_loop(int tid, int lane, int nthreads, uint32_t qn, const uint8_t *tile, const uint16_t *queue,
                         uint16_t *klist, uint32_t *kcount, uint16_t *plane, int t, int n, uint32_t tag) {
    const int l = lane < 0 ? 0 : lane;
    uint32_t ib = (uint32_t)(tid - l);
    if (ib >= qn) return;
    uint32_t ent = queue[ib + (uint32_t)l < qn ? ib + (uint32_t)l : ib];
    while (true) {  // (warp-uniform trip count)
        const bool valid = ib + (uint32_t)l < qn;
        const uint32_t nb = ib + (uint32_t)nthreads;
        const bool more = nb < qn;
        uint32_t ent_next = 0u;
        if (more) ent_next = queue[nb + (uint32_t)l < qn ? nb + (uint32_t)l : nb];  // (in flight during the arithmetic)
        const int rr = (int)(ent >> 9);
        const int j = (int)((ent >> 5) & 15u) * 16 + mask_bit_to_px((int)(ent & 31u));
        FDF_BOUND(rr, SR);
        FDF_BOUND(j - 3, kTileW - 6);  // the ring stays inside the tile row
        const KeypointTest r = test_pixel<MODE, NFIX>(tile + (rr + 3) * kTileW + j, t, n);
        const bool kp = valid && r.kp;
        if (kp) {
            FDF_BOUND(rr * kPlaneW + j - kPlaneLead, SR * kPlaneW);
            FDF_BOUND(j - kPlaneLead, kPlaneW);
            plane[rr * kPlaneW + j - kPlaneLead] = (uint16_t)((tag << 12) | r.score);
        }
#if defined(__CUDA_ARCH__)
        const uint32_t b = __ballot_sync(0xffffffffu, kp);
        if (b != 0u) {
            uint32_t base = 0u;
            if (lane == 0) base = atomicAdd(kcount, (uint32_t)__popc(b));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (kp) {
                FDF_BOUND(base + (uint32_t)__popc(b & ((1u << lane) - 1u)), kQueueCap);
                klist[base + (uint32_t)__popc(b & ((1u << lane) - 1u))] = (uint16_t)((rr << 8) | j);
            }
        }
#else
        if (kp) {
            FDF_BOUND(*kcount, kQueueCap);
            klist[(*kcount)++] = (uint16_t)((rr << 8) | j);
        }
#endif
        if (!more) break;
        ib = nb;
        ent = ent_next;
    }
}

// The loop is instantiated for n = 9 (the reference's and OpenCV's default) and once more for "any n": the run-time
// form costs an indirect branch per candidate.
template <int MODE, int SR>
FDF_HD void phase_b(int tid, int lane, int nthreads, uint32_t qn, const uint8_t *tile, const uint16_t *queue,
                    uint16_t *klist, uint32_t *kcount, uint16_t *plane, int t, int n, uint32_t tag) {
    if (n == 9) phase_b_loop<MODE, SR, 9>(tid, lane, nthreads, qn, tile, queue, klist, kcount, plane, t, n, tag);
    else phase_b_loop<MODE, SR, 0>(tid, lane, nthreads, qn, tile, queue, klist, kcount, plane, t, n, tag);
}

// Dense path (the candidate queue overflowed: very dense content, e.g. noise): every scored pixel of the chunk gets
// the full test straight from the tile, no filter, no queue, no list -- warp `twarp` of `ntwarps` takes scored rows
// twarp, twarp + ntwarps, ..., its lanes the columns.  nms_dense then scans the plane.
template <int MODE, int SR>
FDF_HD void phase_b_dense(int twarp, int lane, int ntwarps, const uint8_t *tile, uint16_t *plane, const ChunkGeo &g,
                          int chunk, int t, int n, uint32_t tag) {
    const RowRange rows = live_rows(g, SR);
    const ColRange cols = scored_cols<MODE>(g.w, chunk);
    const int l0 = lane < 0 ? 0 : lane, lstep = lane < 0 ? 1 : 32;
    for (int rr = rows.lo + twarp; rr < rows.hi; rr += ntwarps)
        for (int j = cols.lo + l0; j < cols.hi; j += lstep) {
            FDF_BOUND(rr, SR);
            FDF_BOUND(j - 3, kTileW - 6);
            FDF_BOUND(j - kPlaneLead, kPlaneW);
            const KeypointTest r = test_pixel<MODE, 0>(tile + (rr + 3) * kTileW + j, t, n);
            if (r.kp) plane[rr * kPlaneW + j - kPlaneLead] = (uint16_t)((tag << 12) | r.score);
        }
}

// ---- NMS: strict maximum over the 8 neighbours (replaces fast_simd.rs:588-616) -------------------
// Only this chunk's own columns and this strip's own rows are emitted; rows 3 and h-4 are scored
// (they act as neighbours) but never emitted (fast_simd.rs:589-596, opencv_compat.rs:238-240).
// Plane cells hold tag << 12 | score with score >= 1.  Tags only grow between two clears of the plane, so a stale
// cell (an earlier chunk's) is smaller than tag << 12, i.e. smaller than any current cell: comparing the raw cells
// is the same as comparing the scores with stale cells read as "no keypoint".
FDF_HD uint32_t max3u(uint32_t a, uint32_t b, uint32_t c) {
#if defined(__CUDA_ARCH__)
    return __vimax3_u32(a, b, c);
#else
    return max(a, max(b, c));
#endif
}

// may the keypoint at (scored row rr, tile column j) be emitted by this chunk at all?
template <int MODE, int SR>
FDF_HD bool nms_emits(int rr, int j, const ChunkGeo &g) {
    const int y = g.ys0 + rr, x = g.xt0 + j;
    if (MODE == NMS_OFF) return x >= g.x0 && x < g.x1;  // (rows and the image border were settled by the filter)
    return !(rr < 1 || rr > SR - 2 || x < g.x0 || x >= g.x1 || y >= g.h - 4);
}

// is the (current) cell pp a strict maximum of its 3x3 neighbourhood?  (branch-free; pp must have a full
// neighbourhood inside the plane)
FDF_HD bool nms_is_max(const uint16_t *pp) {
    const uint32_t a = max3u(pp[-kPlaneW - 1], pp[-kPlaneW], pp[-kPlaneW + 1]);
    const uint32_t b = max3u(pp[kPlaneW - 1], pp[kPlaneW], pp[kPlaneW + 1]);
    const uint32_t c = max3u(pp[-1], pp[1], a);
    return (uint32_t)pp[0] > max(b, c);
}

// staged form of a keypoint: row inside the strip's emitted rows << 16 | image column
template <int MODE>
FDF_HD uint32_t staged_entry(int rr, int j, const ChunkGeo &g) {
    return (uint32_t)((rr - (MODE == NMS_OFF ? 0 : 1)) << 16) | (uint32_t)(g.xt0 + j);
}

// The chunk's keypoint list: every keypoint that survives the NMS (Off mode: every keypoint of the chunk's own
// columns) is written to the staging buffer at base + slot, slots handed out through *scount.  A slot beyond the
// staging buffer's capacity is not written; the function then returns true and the kernel raises the overflow flag
// (the count still includes the keypoint, so the host sees both the needed size and the flag).
template <int MODE, int SR>
FDF_HD bool emit_list(int tid, int nthreads, uint32_t kn, const uint16_t *klist, const uint16_t *plane,
                      uint32_t *scount, unsigned long long base, unsigned long long cap, uint32_t *staging,
                      const ChunkGeo &g) {
    bool dropped = false;
    for (uint32_t i = (uint32_t)tid; i < kn; i += (uint32_t)nthreads) {
        const uint32_t ent = klist[i];
        const int rr = (int)(ent >> 8), j = (int)(ent & 0xffu);
        const bool in = nms_emits<MODE, SR>(rr, j, g);
        bool keep = in;
        // (a keypoint that cannot be emitted is looked up at a harmless cell with a full neighbourhood)
        if (MODE != NMS_OFF) {
            const int cell = in ? rr * kPlaneW + j - kPlaneLead : kPlaneW + 1;
            FDF_BOUND(cell - kPlaneW - 1, SR * kPlaneW);
            FDF_BOUND(cell + kPlaneW + 1, SR * kPlaneW);
            keep = nms_is_max(plane + cell) && in;
        }
        if (keep) {
            const unsigned long long o = base + atomic_add_u32(scount, 1u);
            if (o < cap) staging[o] = staged_entry<MODE>(rr, j, g);
            else dropped = true;
        }
    }
    return dropped;
}

// Dense path: every cell of the plane.  Pass 0 counts the survivors, pass 1 writes them to staging[base + slot]
// with slots handed out through *counter.  Returns true if an entry did not fit the staging buffer.
template <int MODE, int SR>
FDF_HD bool nms_dense(int tid, int nthreads, int pass, const uint16_t *plane, uint32_t *counter, unsigned long long base,
                      unsigned long long cap, uint32_t *staging, const ChunkGeo &g, uint32_t tag) {
    bool dropped = false;
    for (int i = tid; i < SR * kPlaneW; i += nthreads) {
        if (plane[i] < (tag << 12)) continue;
        const int rr = i / kPlaneW, j = i % kPlaneW + kPlaneLead;
        if (!nms_emits<MODE, SR>(rr, j, g)) continue;
        if (MODE != NMS_OFF) {
            FDF_BOUND(i - kPlaneW - 1, SR * kPlaneW);
            FDF_BOUND(i + kPlaneW + 1, SR * kPlaneW);
            if (!nms_is_max(plane + i)) continue;
        }
        const uint32_t slot = atomic_add_u32(counter, 1u);
        if (pass == 1) {
            if (base + slot < cap) staging[base + slot] = staged_entry<MODE>(rr, j, g);
            else dropped = true;
        }
    }
    return dropped;
}

// ---- emission (gather kernel): bit plane -> points, row-major ------------------------------------------
// writes the points of one bit-plane word (row `y`, columns xw .. xw+31) starting at index o
FDF_HD void emit_word(uint32_t m, uint32_t xw, uint32_t y, unsigned long long o, unsigned long long cap, uint2 *out) {
    while (m) {
        const int b = lowest_set_bit(m);
        m &= m - 1u;
        if (o < cap) out[o] = make_uint2(xw + (uint32_t)b, y);
        o++;
    }
}

}  // namespace fdf
