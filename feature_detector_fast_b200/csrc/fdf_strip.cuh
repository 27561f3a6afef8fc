// fdf_strip.cuh -- the per-thread bodies of the detection kernel's phases.
//
// They are `__host__ __device__` and take the thread index as an argument: fdf_kernels.cu calls
// them with threadIdx.x between its barriers, and tests/host/strip_emulator.cpp runs the very same
// code thread by thread on the CPU (TMA replaced by a zero-filled copy) so that the tiling,
// halo, validity and NMS-row rules are checked against the oracle without a GPU.
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
    int ww;     // bit-plane words per row
};

template <int MODE>
FDF_HD ChunkGeo make_geo(int w, int h, int ww, int strip, int chunk, int sr) {
    ChunkGeo g;
    g.w = w;
    g.h = h;
    g.ww = ww;
    g.y0 = first_out_row(MODE) + strip * out_rows(MODE, sr);
    g.ys0 = g.y0 - (MODE == NMS_OFF ? 0 : 1);
    g.xt0 = chunk * kChunkW - kTileLead;
    g.x0 = g.xt0 + kLeftHalo;
    g.x1 = (chunk == chunks_per_row(w) - 1) ? w - 3 : g.x0 + kChunkW;
    return g;
}

// ---- phase A: dense filter, 16 centres per thread and row (replaces fast_simd.rs:368-520) -------
// Thread t handles the 16-pixel group t & 15 of scored rows (t >> 4) + 16k inside [row_lo, row_hi) and
// pushes (scored row << 8 | tile column) of each centre that passes the necessary-condition filter to the
// CTA's candidate queue.  Column validity (image border, chunk halo) is checked in phase B.  When the queue
// is full the entries are dropped but still counted: *qcount > kQueueCap tells the caller to redo the chunk
// in row groups.
template <int MODE, int SR>
FDF_HD void phase_a(int tid, const uint8_t *tile, uint16_t *queue, uint32_t *qcount, const ChunkGeo &g,
                    uint32_t kbias, int row_lo, int row_hi) {
    constexpr int IT = SR / (kComputeThreads / 16);  // rows per thread
    const int q = tid & 15;                          // which 16-pixel group of the 256-wide tile row
    const int r0 = tid >> 4;                         // first scored row of this thread
    // step 1, straight-line for all of the thread's rows (independent work the scheduler can overlap):
    // bit (8*b + k) of gm[it] <=> byte b of word k of the group passed the filter
    uint32_t gm[IT];
#pragma unroll
    for (int it = 0; it < IT; it++) {
        const int rr = r0 + it * (kComputeThreads / 16);
        const int y = g.ys0 + rr;
        const bool live = y >= 3 && y < g.h - 3 && rr >= row_lo && rr < row_hi;  // fast_simd.rs:342
        const uint8_t *rowp = tile + (rr + 3) * kTileW + q * 16;
        const uint4 C = *reinterpret_cast<const uint4 *>(rowp);
        const uint4 N = *reinterpret_cast<const uint4 *>(rowp - 3 * kTileW);
        const uint4 S = *reinterpret_cast<const uint4 *>(rowp + 3 * kTileW);
        const uint32_t cl = q > 0 ? *reinterpret_cast<const uint32_t *>(rowp - 4) : 0u;
        const uint32_t cr = q < 15 ? *reinterpret_cast<const uint32_t *>(rowp + 16) : 0u;
        // east = pixel x+3, west = pixel x-3 of the same row: byte-shifted views of the row words
        const uint32_t e0 = byte_perm(C.x, C.y, 0x6543), e1 = byte_perm(C.y, C.z, 0x6543);
        const uint32_t e2 = byte_perm(C.z, C.w, 0x6543), e3 = byte_perm(C.w, cr, 0x6543);
        const uint32_t w0 = byte_perm(cl, C.x, 0x4321), w1 = byte_perm(C.x, C.y, 0x4321);
        const uint32_t w2 = byte_perm(C.y, C.z, 0x4321), w3 = byte_perm(C.z, C.w, 0x4321);
        const uint32_t f0 = filter4(C.x, N.x, S.x, e0, w0, kbias, 0x80808080u);
        const uint32_t f1 = filter4(C.y, N.y, S.y, e1, w1, kbias, 0x80808080u);
        const uint32_t f2 = filter4(C.z, N.z, S.z, e2, w2, kbias, 0x80808080u);
        const uint32_t f3 = filter4(C.w, N.w, S.w, e3, w3, kbias, 0x80808080u);
        // each f has only bit 7 of its bytes set
        gm[it] = live ? ((f0 >> 7) | (f1 >> 6) | (f2 >> 5) | (f3 >> 4)) : 0u;
    }
    // step 2: push the survivors
#pragma unroll
    for (int it = 0; it < IT; it++) {
        uint32_t m = gm[it];
        if (m != 0u) {
            const int rr = r0 + it * (kComputeThreads / 16);
            const uint32_t cnt = (uint32_t)popc32(m);
            uint32_t slot = atomic_add_u32(qcount, cnt);
            if (slot + cnt <= (uint32_t)kQueueCap) {
                const uint32_t ent0 = (uint32_t)((rr << 8) | (q * 16));
                while (m) {
                    const uint32_t p = (uint32_t)lowest_set_bit(m);
                    m &= m - 1u;
                    queue[slot++] = (uint16_t)(ent0 + ((p & 7u) << 2) + (p >> 3));
                }
            }
        }
    }
}

// ---- phase B: exact segment test (+ score) per candidate (replaces fast_simd.rs:115-297, 623-749)
// One thread per queue entry.  Off mode: sets the keypoint's bit in the strip bit plane.  NMS modes: writes
// (tag << 12 | score) into the score plane at (scored row, tile column) and appends the entry to the
// chunk's keypoint list (entries beyond kKlistCap are only counted: the caller then runs the dense NMS).
template <int MODE, int SR>
FDF_HD void phase_b(int tid, uint32_t qn, const uint8_t *tile, const uint16_t *queue, uint16_t *plane,
                    uint16_t *klist, uint32_t *kcount, uint32_t *bits, const ChunkGeo &g, int t, int n,
                    uint32_t tag) {
    constexpr int HS = (MODE == NMS_OFF) ? 0 : 1;
    // scored columns: the chunk's own columns plus the NMS score halo, inside the image's centre range
    const int xlo = max(3, g.x0 - HS), xhi = min(g.w - 3, g.x1 + HS);
    for (uint32_t i = (uint32_t)tid; i < qn; i += (uint32_t)kComputeThreads) {
        const uint32_t ent = queue[i];
        const int rr = (int)(ent >> 8), j = (int)(ent & 0xffu);
        const int x = g.xt0 + j;
        if (x < xlo || x >= xhi) continue;  // fast_simd.rs:369-371, 559-562
        const uint8_t *pc = tile + (rr + 3) * kTileW + j;
        const int cv = pc[0];
        int ring[16];
#pragma unroll
        for (int k = 0; k < 16; k++) ring[k] = pc[FDF_RING_DY(k) * kTileW + FDF_RING_DX(k)];
        const RingMasks rm = ring_masks(cv, ring, t);
        const bool arc_bright = has_arc(rm.bright, n);
        const bool arc_dark = has_arc(rm.dark, n);
        if (arc_bright || arc_dark) {
            if (MODE == NMS_OFF) {
                atomic_or_u32(&bits[rr * g.ww + (x >> 5)], 1u << (x & 31));
            } else {
                const uint32_t sc = (MODE == NMS_MAX_THRESHOLD) ? score_max_threshold(cv, ring, n, arc_bright)
                                                                : score_sum_abs(cv, ring, t);  // <= 4080 < 2^12
                plane[rr * kTileW + j] = (uint16_t)((tag << 12) | sc);
                const uint32_t k = atomic_add_u32(kcount, 1u);
                if (k < (uint32_t)kKlistCap) klist[k] = (uint16_t)ent;
            }
        }
    }
}

// ---- NMS: strict maximum over the 8 neighbours (replaces fast_simd.rs:588-616) -------------------
// Only this chunk's own columns and this strip's own rows are emitted; rows 3 and h-4 are scored
// (they act as neighbours) but never emitted (fast_simd.rs:589-596, opencv_compat.rs:238-240).
// A plane entry belongs to the current chunk iff its tag does; anything else is stale = "no keypoint".
FDF_HD uint32_t live_score(uint32_t v, uint32_t tag_floor) { return v >= tag_floor ? (v & 0xfffu) : 0u; }

template <int MODE, int SR>
FDF_HD void nms_one(int rr, int j, const uint16_t *plane, uint32_t *bits, const ChunkGeo &g, uint32_t floor) {
    const int y = g.ys0 + rr, x = g.xt0 + j;
    if (rr < 1 || rr > SR - 2 || x < g.x0 || x >= g.x1 || y >= g.h - 4) return;
    const uint16_t *pp = plane + rr * kTileW + j;
    const uint32_t s = live_score(pp[0], floor);
    if (s == 0u) return;
    const bool keep = s > live_score(pp[-kTileW - 1], floor) && s > live_score(pp[-kTileW], floor) &&
                      s > live_score(pp[-kTileW + 1], floor) && s > live_score(pp[-1], floor) &&
                      s > live_score(pp[1], floor) && s > live_score(pp[kTileW - 1], floor) &&
                      s > live_score(pp[kTileW], floor) && s > live_score(pp[kTileW + 1], floor);
    if (keep) atomic_or_u32(&bits[(rr - 1) * g.ww + (x >> 5)], 1u << (x & 31));
}

// the chunk's keypoint list (the common case)
template <int MODE, int SR>
FDF_HD void nms_list(int tid, uint32_t kn, const uint16_t *klist, const uint16_t *plane, uint32_t *bits,
                     const ChunkGeo &g, uint32_t tag) {
    for (uint32_t i = (uint32_t)tid; i < kn; i += (uint32_t)kComputeThreads) {
        const uint32_t ent = klist[i];
        nms_one<MODE, SR>((int)(ent >> 8), (int)(ent & 0xffu), plane, bits, g, tag << 12);
    }
}

// every cell of the plane (only when the list overflowed: very dense content)
template <int MODE, int SR>
FDF_HD void nms_dense(int tid, const uint16_t *plane, uint32_t *bits, const ChunkGeo &g, uint32_t tag) {
    for (int i = tid; i < SR * kTileW; i += kComputeThreads) {
        if (plane[i] >= (tag << 12)) nms_one<MODE, SR>(i / kTileW, i % kTileW, plane, bits, g, tag << 12);
    }
}

// ---- emission: bit plane -> points, row-major -----------------------------------------------------
// The strip's bit plane (out_rows x ww words, in 128-bit units) is cut into one contiguous range per warp;
// a warp walks its range 32 units at a time: one unit per lane, a warp prefix sum gives every lane its offset.
struct EmitRange {
    int begin, end;  // unit indices
};

FDF_HD EmitRange emit_range(int warp, int nunits) {
    const int upw = (nunits + kThreads / 32 - 1) / (kThreads / 32);
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
        if (o < cap) {
            out[o].x = xw + (uint32_t)b;
            out[o].y = y;
        }
        o++;
    }
}

}  // namespace fdf
