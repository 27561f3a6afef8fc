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

FDF_HD void atomic_or_u32(uint32_t *p, uint32_t v) {
#if defined(__CUDA_ARCH__)
    atomicOr(p, v);
#else
    *p |= v;
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
// image's centre range [3, w-3).  The table has one word per 4 tile columns with 0x80 in every valid byte; it
// depends on the chunk only through "first / middle / last chunk of the row", so the kernel builds the
// three variants once (kVtabWords words each).
constexpr int kVtabWords = kTileW / 4;

template <int MODE>
FDF_HD uint32_t valid_word(int w, int chunk, int word) {
    constexpr int HS = (MODE == NMS_OFF) ? 0 : 1;
    const int xt0 = chunk * kChunkW - kTileLead, x0 = xt0 + kLeftHalo;
    const int x1 = (chunk == chunks_per_row(w) - 1) ? w - 3 : x0 + kChunkW;
    const int xlo = max(3, x0 - HS), xhi = min(w - 3, x1 + HS);
    uint32_t v = 0u;
    for (int b = 0; b < 4; b++) {
        const int x = xt0 + 4 * word + b;
        if (x >= xlo && x < xhi) v |= 0x80u << (8 * b);
    }
    return v;
}

// variant of chunk c: 2 = last chunk of the row (also when it is the only one), 0 = first, 1 = any other
FDF_HD int vtab_variant(int c, int nc) { return c == nc - 1 ? 2 : (c == 0 ? 0 : 1); }
FDF_HD int vtab_chunk(int variant, int nc) { return variant == 2 ? nc - 1 : variant; }

// ---- phase A: two-stage dense filter (replaces fast_simd.rs:368-520) -------------------------------------
// Stage 1 looks at every scored pixel, 16 per lane and row, north/south pair only (~12 % of the pixels of
// natural content pass, ~19 % of the 16-pixel groups).  Groups with a survivor are compacted into the warp's
// own queue (ballot + rank, no atomics, no block barrier); stage 2 then runs the full two-pair filter with one
// lane per queued group and pushes every surviving centre (~2 % of the pixels) to the CTA's candidate queue.
//
// Warp queue entry: scored row << 4 | group.  Candidate entry: scored row << 9 | group << 5 | mask bit.

// scored rows [lo, hi) of a strip that can hold a centre at all (fast_simd.rs:342: image rows 3 .. h-4),
// intersected with the row range the caller asks for
struct RowRange {
    int lo, hi;
};
FDF_HD RowRange live_rows(const ChunkGeo &g, int row_lo, int row_hi) {
    RowRange r;
    r.lo = max(row_lo, 3 - g.ys0);
    r.hi = min(row_hi, g.h - 3 - g.ys0);
    return r;
}

// Stage 1 for one lane: the 16-pixel group q of BH consecutive scored rows rr0 .. rr0+BH-1.
//   v bit i: row rr0 + i has a centre whose north or south ring pixel differs from it by more than t;
//   h bit i: row rr0 + i has a pixel y in the group with |p(y + 3) - p(y)| > t.
// A candidate centre x needs max(|E-c|, |W-c|) > t, i.e. the horizontal difference at y = x or at y = x - 3, and
// y = x - 3 may lie in the group to the left: the caller ORs the left neighbour's h into this group's.  With
// -DFDF_STAGE1_H a (row, group) goes on to stage 2 only if v and (h or left h): ~9 % of them instead of the ~18 %
// that pass v alone.  Without it (default, see below) h is all ones and v alone decides.
// The rows are walked top to bottom with everything kept in registers: every tile row is loaded once (LDS.128)
// and the vertical difference D(y) = |p(y) - p(y-3)| is computed once and used twice (as the north difference of
// centre y and the south difference of centre y - 3).
struct Stage1Masks {
    uint32_t v, h;
};

template <int BH>
FDF_HD Stage1Masks stage1_band(const uint8_t *tile, int rr0, int q, uint32_t kbias) {
    const uint8_t *p = tile + rr0 * kTileW + q * 16;  // tile row rr0 is the north ring row of scored row rr0
    Px16 row[BH + 6], d[BH + 3];
    Stage1Masks m;
    m.v = m.h = 0u;
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
            const uint32_t o = exceeds4(x, kbias);
            uint32_t oh = 0u;
            if ((o & 0x80808080u) != 0u) m.v |= 1u << (i - 6);
#if !defined(FDF_STAGE1_H)  // the horizontal group test below is a measured loss (1.077 vs 1.039 ms per 256 frames: its
                            // dense work costs more than the stage-2 entries it saves), so it is off by default
            m.h = 0xffffffffu;
            continue;
#endif
            // the word right of the group = the first word of lane + 1's group (same rows); for q = 15 it is some
            // other pixel word, which can only set h where it need not be set
#if defined(__CUDA_ARCH__)
            const uint32_t wr = __shfl_down_sync(0xffffffffu, row[i - 3].w[0], 1);  // (no shared-memory wavefronts)
#else
            const uint32_t wr = *reinterpret_cast<const uint32_t *>(p + (i - 3) * kTileW + 16);
#endif
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t east = byte_perm(row[i - 3].w[k], k < 3 ? row[i - 3].w[k + 1] : wr, 0x6543u);
                oh |= exceeds4(absdiff4(east, row[i - 3].w[k]), kbias);
            }
            if ((oh & 0x80808080u) != 0u) m.h |= 1u << (i - 6);
        }
    }
    return m;
}

// bit i set iff scored row rr0 + i may hold a centre at all (fast_simd.rs:342: image rows 3 .. h-4) and lies in
// the row range the caller asks for
FDF_HD uint32_t live_mask(const ChunkGeo &g, int rr0, int bh, int row_lo, int row_hi) {
    const RowRange live = live_rows(g, row_lo, row_hi);
    uint32_t m = 0u;
    for (int i = 0; i < bh; i++)
        if (rr0 + i >= live.lo && rr0 + i < live.hi) m |= 1u << i;
    return m;
}

// stage 2 for one queued group: the candidate mask of its 16 centres (bit layout: candidate_mask16)
FDF_HD uint32_t stage2_mask(uint32_t e, const uint8_t *tile, const uint32_t *vtab, uint32_t all_valid_inside,
                            uint32_t kbias) {
    const int rr = (int)(e >> 4), q = (int)(e & 15u);
    const uint8_t *rowp = tile + (rr + 3) * kTileW + q * 16;
    // the words left / right of the group: for q = 0 / q = 15 they belong to the neighbouring tile row, which
    // only reaches centres the validity table excludes (tile columns 0..2 and 253..255)
#if defined(FDF_STAGE2_LDS64)  // timing experiments (measured slower: 1.079 vs 1.035 ms)
    const uint32_t cl = reinterpret_cast<const uint2 *>(rowp - 8)->y;
    const uint32_t cr = reinterpret_cast<const uint2 *>(rowp + 16)->x;
#else
    const uint32_t cl = *reinterpret_cast<const uint32_t *>(rowp - 4);
    const uint32_t cr = *reinterpret_cast<const uint32_t *>(rowp + 16);
#endif
#if defined(FDF_STAGE2_VTAB_PRED)
    // validity: only the first two and the last group of a tile row (and the row's last chunk) have excluded centres
    uint32_t valid[4] = {0x80808080u, 0x80808080u, 0x80808080u, 0x80808080u};
    if (q <= 1 || q == 15 || all_valid_inside == 0u) {
        const uint4 vv = *reinterpret_cast<const uint4 *>(vtab + 4 * q);
        valid[0] = vv.x, valid[1] = vv.y, valid[2] = vv.z, valid[3] = vv.w;
    }
#else
    (void)all_valid_inside;
    const uint4 vv = *reinterpret_cast<const uint4 *>(vtab + 4 * q);
    const uint32_t valid[4] = {vv.x, vv.y, vv.z, vv.w};
#endif
    return candidate_mask16(load16(rowp), load16(rowp - 3 * kTileW), load16(rowp + 3 * kTileW), cl, cr, valid, kbias);
}

// one queue entry per set bit of the group's candidate mask, at queue[slot ...]
FDF_HD void push_candidates(uint32_t e, uint32_t m, uint16_t *queue, uint32_t slot) {
    const uint32_t base = (e >> 4) << 9 | (e & 15u) << 5;
    uint16_t *out = queue + slot;
#if !defined(FDF_PUSH_STRAIGHT)  // bit walk (default; the predicated straight-line form below measured 0.7 % slower)
    while (m != 0u) {
        const uint32_t p = (uint32_t)highest_set_bit(m);
        m ^= 1u << p;
        *out++ = (uint16_t)(base + p);
    }
    return;
#endif
#pragma unroll
    for (int p = 0; p < 32; p++) {
        if (((0xf0f0f0f0u >> p) & 1u) == 0u) continue;  // candidate_mask16 uses bits 8b + 7 - k
        if ((m >> p) & 1u) {
            *out = (uint16_t)(base + (uint32_t)p);
            out++;
        }
    }
}

// One warp's phase A; NW warps share a chunk.  On the device the 32 lanes run it together (lane = threadIdx.x & 31);
// on the host the emulator calls it once per warp and the lane loops below run sequentially, in the same order as
// the ballot ranks.  wq is the warp's private queue (kWarpQueueCap entries).
// Lane l of warp v handles 16-pixel group q = l & 15 of the BH = SR / (2 NW) scored rows starting at
// (2 v + (l >> 4)) * BH.
template <int MODE, int SR, int NW>
FDF_HD void phase_a_warp(int warp, int lane_or_minus1, const uint8_t *tile, uint16_t *wq, const uint32_t *vtab,
                         int variant, uint16_t *queue, uint32_t *qcount, const ChunkGeo &g, uint32_t kbias, int row_lo,
                         int row_hi, long long *trace = nullptr) {
    constexpr int BH = SR / (2 * NW);
    static_assert(BH * 32 * NW <= 4 * kWarpQueueCap, "a warp queue must hold every group stage 1 looks at");
    uint32_t n = 0u;  // entries in wq (warp-uniform)
    // groups 2 .. 14 of a chunk are all-valid unless it is the last chunk of its row (variant 2)
    const uint32_t inside = variant == 2 ? 0u : 1u;
#if defined(__CUDA_ARCH__)
    const int lane = lane_or_minus1;
    const int q = lane & 15, rr0 = (2 * warp + (lane >> 4)) * BH;
    const uint32_t lt = (1u << lane) - 1u;
    const Stage1Masks s1 = stage1_band<BH>(tile, rr0, q, kbias);
    uint32_t hl = __shfl_up_sync(0xffffffffu, s1.h, 1);  // the group to the left, same rows
    if (q == 0) hl = 0u;  // (tile columns 0 .. 15: the centres that count start at column 11, their x - 3 is in the group)
    const uint32_t need = s1.v & (s1.h | hl) & live_mask(g, rr0, BH, row_lo, row_hi);
    const uint32_t ent = (uint32_t)((rr0 << 4) | q);
#if !defined(FDF_WQ_SCAN)  // one ballot per row (default)
#pragma unroll
    for (int i = 0; i < BH; i++) {
        const bool mine = (need >> i) & 1u;
        const uint32_t b = __ballot_sync(0xffffffffu, mine);
        if (mine) wq[n + (uint32_t)__popc(b & lt)] = (uint16_t)(ent + (uint32_t)(i << 4));
        n += (uint32_t)__popc(b);
    }
#else
    {   // timing experiment: warp prefix sum of the lanes' row counts, then every lane writes its own rows.  Fewer
        // instructions (~45 instead of ~80 per lane) but measured 16 % SLOWER (1.193 vs 1.029 ms per 256 frames): the
        // five dependent shuffles sit on the filter warps' critical path and the lane-major entry order makes the
        // stage-2 row loads of a quarter-warp hit the same banks.
        (void)lt;
        const uint32_t cnt = (uint32_t)__popc(need);
        uint32_t incl = cnt;
#pragma unroll
        for (int dd = 1; dd < 32; dd <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, dd);
            if (lane >= dd) incl += v;
        }
        n = __shfl_sync(0xffffffffu, incl, 31);
        uint16_t *out = wq + (incl - cnt);
#pragma unroll
        for (int i = 0; i < BH; i++) {
            if ((need >> i) & 1u) {
                *out = (uint16_t)(ent + (uint32_t)(i << 4));
                out++;
            }
        }
    }
#endif
    __syncwarp();
#if defined(FDF_TRACE)
    if (trace != nullptr) trace[9] = clock64();  // stage 1 done
#endif
#if !(defined(FDF_ABLATE) && (FDF_ABLATE & 32))  // timing experiment: stage 1 only
    // stage 2, one lane per queued group.  When the queue is full the entries are dropped but still counted:
    // *qcount > kQueueCap tells the test warps to redo the chunk in row groups.
#if defined(FDF_STAGE2_X2)  // timing experiment: two queued groups per lane and step (interleaved dependency chains)
    for (uint32_t i = (uint32_t)lane; i < n; i += 64u) {
        const bool two = i + 32u < n;
        const uint32_t e0 = wq[i], e1 = wq[two ? i + 32u : i];
        const uint32_t m0 = stage2_mask(e0, tile, vtab, inside, kbias);
        uint32_t m1 = stage2_mask(e1, tile, vtab, inside, kbias);
        if (!two) m1 = 0u;
        const uint32_t c0 = (uint32_t)__popc(m0), c1 = (uint32_t)__popc(m1);
        if (c0 + c1 != 0u) {
            const uint32_t slot = atomicAdd(qcount, c0 + c1);
            if (slot + c0 + c1 <= (uint32_t)kQueueCap) {
                push_candidates(e0, m0, queue, slot);
                push_candidates(e1, m1, queue, slot + c0);
            }
        }
    }
#else
    for (uint32_t i = (uint32_t)lane; i < n; i += 32u) {
        const uint32_t e = wq[i];
        const uint32_t m = stage2_mask(e, tile, vtab, inside, kbias);
#if defined(FDF_ABLATE) && (FDF_ABLATE & 64)  // timing experiment: stage 2 without the candidate push
        if (m == 0xdeadbeefu) queue[0] = (uint16_t)m;
        continue;
#endif
        if (m != 0u) {
            const uint32_t cnt = (uint32_t)__popc(m);
            const uint32_t slot = atomicAdd(qcount, cnt);
            if (slot + cnt <= (uint32_t)kQueueCap) push_candidates(e, m, queue, slot);
        }
    }
#endif
#endif
#else
    (void)lane_or_minus1;
    (void)trace;
    uint32_t hprev = 0u;
    for (int lane = 0; lane < 32; lane++) {
        const int q = lane & 15, rr0 = (2 * warp + (lane >> 4)) * BH;
        const Stage1Masks s1 = stage1_band<BH>(tile, rr0, q, kbias);
        const uint32_t need = s1.v & (s1.h | (q == 0 ? 0u : hprev)) & live_mask(g, rr0, BH, row_lo, row_hi);
        hprev = s1.h;
        for (int i = 0; i < BH; i++)
            if ((need >> i) & 1u) wq[n++] = (uint16_t)(((rr0 + i) << 4) | q);
    }
    for (uint32_t i = 0; i < n; i++) {
        const uint32_t m = stage2_mask(wq[i], tile, vtab, inside, kbias);
        const uint32_t cnt = (uint32_t)popc32(m);
        if (cnt != 0u) {
            const uint32_t slot = atomic_add_u32(qcount, cnt);
            if (slot + cnt <= (uint32_t)kQueueCap) push_candidates(wq[i], m, queue, slot);
        }
    }
#endif
}

// ---- phase B: exact segment test (+ score) per candidate (replaces fast_simd.rs:115-297, 623-749)
// One thread per queue entry: 16 ring bytes + the centre from the tile, one dual word per ring pixel (fdf_core.cuh),
// best window -> keypoint yes / no and the MaxThreshold score in the same ~50 instructions.  Every keypoint writes
// (tag << 12 | score) into the score plane at (scored row, tile column - kPlaneLead) -- in Off mode the score is 1
// and only the dense fallback reads it -- and is appended to the chunk's keypoint list as scored row << 8 | tile
// column (one ballot per warp step, one shared atomic per warp step that found a keypoint).  klist == nullptr
// (dense fallback): no list.
// On the device the 32 lanes of a warp call it together (lane >= 0); the host emulator calls it once per thread
// with lane = -1.
// `tile_done()` is called exactly once per call, warp-uniformly, as soon as this thread has read everything it needs
// from the tile and the candidate queue (after the loads of its last step): the kernel uses it to request the next
// tile before the arithmetic of the last step instead of after it.
struct NoTileDone {
    FDF_HD void operator()() const {}
};

template <int MODE, int SR, int U = 1, class TileDone = NoTileDone>
FDF_HD void phase_b(int tid, int lane, int nthreads, uint32_t qn, const uint8_t *tile, const uint16_t *queue,
                    uint16_t *klist, uint32_t *kcount, uint16_t *plane, int t, int n, uint32_t tag,
                    TileDone tile_done = TileDone()) {
    const int l = lane < 0 ? 0 : lane;
    bool told = false;
    // U queue entries per thread and step, as separate load / arithmetic / store sections, so that their dependency
    // chains interleave (the test warps are few; a single chain leaves them waiting on shared-memory latency)
    for (uint32_t ib = (uint32_t)(tid - l); ib < qn; ib += (uint32_t)(U * nthreads)) {  // (warp-uniform trip count)
        RingDual ring[U];
        uint32_t pos[U];
        bool valid[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint32_t i = ib + (uint32_t)(u * nthreads) + (uint32_t)l;
            valid[u] = i < qn;
            const uint32_t ent = queue[valid[u] ? i : ib];
            const int rr = (int)(ent >> 9);
            const int j = (int)((ent >> 5) & 15u) * 16 + mask_bit_to_px((int)(ent & 31u));
            pos[u] = (uint32_t)((rr << 8) | j);
            const uint8_t *pc = tile + (rr + 3) * kTileW + j;
            const uint32_t bias = dual_bias((int)pc[0]);
#pragma unroll
            for (int k = 0; k < 16; k++)
                ring[u].w[k] = dual_word((uint32_t)pc[FDF_RING_DY(k) * kTileW + FDF_RING_DX(k)], bias);
        }
        if (ib + (uint32_t)(U * nthreads) >= qn) {  // (warp-uniform) the last step: the ring words are in registers
            tile_done();
            told = true;
        }
        bool kp[U];
        uint32_t sc[U];
#if defined(FDF_EXP_LOADS) || defined(FDF_EXP_ALU)  // sensitivity experiments: extra shared-memory loads / extra logic-pipe work
        uint32_t dummy = 0u;
#endif
#if defined(FDF_EXP_LOADS)
        {
            const uint32_t ent = queue[valid[0] ? ib + (uint32_t)l : ib];
            const uint8_t *pc = tile + ((int)(ent >> 9) + 3) * kTileW + (int)((ent >> 5) & 15u) * 16 + mask_bit_to_px((int)(ent & 31u));
#pragma unroll
            for (int k = 0; k < 16; k++) dummy += pc[FDF_RING_DY(k) * kTileW + FDF_RING_DX((k + 5) & 15)];
        }
#endif
#if defined(FDF_EXP_ALU)
        {
#if FDF_EXP_ALU == 1   // one dependent chain of 48 logic-pipe instructions (+ 48 XORs)
            uint32_t x = ring[0].w[0];
#pragma unroll
            for (int k = 0; k < 48; k++) x = min3_u16x2(x ^ ring[0].w[k & 15], ring[0].w[(k + 3) & 15], ring[0].w[(k + 7) & 15]);
            dummy += x;
#elif FDF_EXP_ALU == 2  // 96 logic-pipe instructions in 16 independent chains
            uint32_t x[16];
#pragma unroll
            for (int k = 0; k < 16; k++) x[k] = ring[0].w[k];
#pragma unroll
            for (int r = 0; r < 6; r++)
#pragma unroll
                for (int k = 0; k < 16; k++) x[k] = min3_u16x2(x[k] ^ ring[0].w[(k + r) & 15], ring[0].w[(k + 3) & 15], ring[0].w[(k + 7 + r) & 15]);
#pragma unroll
            for (int k = 0; k < 16; k++) dummy ^= x[k];
#else                   // 96 multiply-adds in 16 independent chains
            uint32_t x[16];
#pragma unroll
            for (int k = 0; k < 16; k++) x[k] = ring[0].w[k];
#pragma unroll
            for (int r = 0; r < 6; r++)
#pragma unroll
                for (int k = 0; k < 16; k++) x[k] = x[k] * ring[0].w[(k + r + 1) & 15] + ring[0].w[(k + 7 + r) & 15];
#pragma unroll
            for (int k = 0; k < 16; k++) dummy ^= x[k];
#endif
        }
#endif
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint32_t best = best_of_lanes(best_window(ring[u], n));
            kp[u] = valid[u] && best > (uint32_t)(256 + t);
#if defined(FDF_EXP_LOADS) || defined(FDF_EXP_ALU)
            if (dummy == 0x7654321u) kp[u] = !kp[u];
#endif
            sc[u] = 1u;  // Off mode: the plane only records "keypoint here" (used by the dense fallback)
            if (MODE == NMS_MAX_THRESHOLD) sc[u] = best - 256u;                     // (garbage unless kp)
            if (MODE == NMS_SUM_ABSOLUTE) sc[u] = score_sum_abs_dual(ring[u], t);  // <= 4080 < 2^12
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (kp[u]) plane[(pos[u] >> 8) * kPlaneW + (pos[u] & 0xffu) - kPlaneLead] = (uint16_t)((tag << 12) | sc[u]);
#if defined(__CUDA_ARCH__)
            if (klist != nullptr) {
                const uint32_t b = __ballot_sync(0xffffffffu, kp[u]);
                if (b != 0u) {
                    uint32_t base = 0u;
                    if (lane == 0) base = atomicAdd(kcount, (uint32_t)__popc(b));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (kp[u]) klist[base + (uint32_t)__popc(b & ((1u << lane) - 1u))] = (uint16_t)pos[u];
                }
            }
#else
            if (klist != nullptr && kp[u]) klist[(*kcount)++] = (uint16_t)pos[u];
#endif
        }
    }
    if (!told) tile_done();  // (no step at all)
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
// columns) is written to the staging buffer at base + slot, slots handed out through *scount.
template <int MODE, int SR>
FDF_HD void emit_list(int tid, int nthreads, uint32_t kn, const uint16_t *klist, const uint16_t *plane,
                      uint32_t *scount, unsigned long long base, unsigned long long cap, uint32_t *staging,
                      const ChunkGeo &g) {
    for (uint32_t i = (uint32_t)tid; i < kn; i += (uint32_t)nthreads) {
        const uint32_t ent = klist[i];
        const int rr = (int)(ent >> 8), j = (int)(ent & 0xffu);
        const bool in = nms_emits<MODE, SR>(rr, j, g);
        bool keep = in;
        // (a keypoint that cannot be emitted is looked up at a harmless cell with a full neighbourhood)
        if (MODE != NMS_OFF) keep = nms_is_max(plane + (in ? rr * kPlaneW + j - kPlaneLead : kPlaneW + 1)) && in;
        if (keep) {
            const unsigned long long o = base + atomic_add_u32(scount, 1u);
            if (o < cap) staging[o] = staged_entry<MODE>(rr, j, g);
        }
    }
}

// Dense fallback (queue overflow: very dense content): every cell of the plane.  Pass 0 counts the
// survivors, pass 1 writes them to staging[base + slot] with slots handed out through *slot_counter.
template <int MODE, int SR>
FDF_HD void nms_dense(int tid, int nthreads, int pass, const uint16_t *plane, uint32_t *counter, unsigned long long base,
                      unsigned long long cap, uint32_t *staging, const ChunkGeo &g, uint32_t tag) {
    for (int i = tid; i < SR * kPlaneW; i += nthreads) {
        if (plane[i] < (tag << 12)) continue;
        const int rr = i / kPlaneW, j = i % kPlaneW + kPlaneLead;
        if (!nms_emits<MODE, SR>(rr, j, g)) continue;
        if (MODE != NMS_OFF && !nms_is_max(plane + i)) continue;
        const uint32_t slot = atomic_add_u32(counter, 1u);
        if (pass == 1 && base + slot < cap) staging[base + slot] = staged_entry<MODE>(rr, j, g);
    }
}

// ---- emission (gather kernel): bit plane -> points, row-major ------------------------------------------
// The strip's bit plane (out_rows x ww words, in 128-bit units) is cut into one contiguous range per warp;
// a warp walks its range 32 units at a time: one unit per lane, a warp prefix sum gives every lane its offset.
struct EmitRange {
    int begin, end;  // unit indices
};

FDF_HD EmitRange emit_range(int warp, int nunits) {
    const int upw = (nunits + kGatherThreads / 32 - 1) / (kGatherThreads / 32);
    EmitRange r;
    r.begin = min(warp * upw, nunits);
    r.end = min(r.begin + upw, nunits);
    return r;
}

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
