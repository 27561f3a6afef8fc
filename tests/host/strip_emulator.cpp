// strip_emulator.cpp -- runs the detection kernel's per-thread phase bodies (fdf_strip.cuh, the
// same source the GPU executes) thread by thread on the CPU, with the TMA load replaced by a
// zero-filled copy and the decoupled look-back replaced by a running offset.  TEST CODE ONLY:
// it exists so the tiling / halo / validity / NMS-row logic and the SWAR arithmetic can be checked
// against the oracle in the CPU-only test tier.  It is not a CPU fallback: the product library
// never contains or calls it.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>
#include <vector>

#include "../../feature_detector_fast_b200/csrc/fdf_strip.cuh"
#include "../../oracle/fdf_oracle.h"
#include "second_impl.h"

namespace {

using namespace fdf;

// Mirrors what the kernel does with one chunk: stage 1 by every filter warp, stage 2 by the filter threads over all
// warps' entries, phase B by the test threads (or the dense path over every scored pixel when the candidate queue
// overflowed), NMS pass over the chunk's keypoint list (or over the whole plane in the dense path), then the chunk's
// surviving keypoints go to the staging buffer as one unordered run.  At the end of a strip the gather kernel's part
// follows: runs -> bit plane -> row-major points.
template <int MODE, int SR>
int64_t emulate(const uint8_t *img, int w, int h, int pitch, int t, int n, uint2 *out, size_t cap, int *fallbacks) {
    constexpr int OUT_R = out_rows(MODE, SR);
    constexpr int TR = tile_rows(SR);
    const long long rows = (long long)h - 2 * first_out_row(MODE);
    if (w < 7 || h < 7 || rows <= 0) return 0;
    const int S = (int)((rows + OUT_R - 1) / OUT_R);
    const int NC = chunks_per_row(w);
    const int WW = (w + 31) / 32;
    alignas(16) static uint8_t tile[tile_rows(64) * kTileW];
    std::vector<uint16_t> plane((size_t)SR * kPlaneW), queue(kQueueCap), klist(kQueueCap);
    std::vector<uint8_t> ents((size_t)kFilterWarps * kWarpQueueCap);
    alignas(16) uint32_t vtab[3][kVtabWords];  // validity tables: first / middle / last chunk of a row
    for (int v = 0; v < 3; v++)
        for (int i = 0; i < kVtabWords; i++) vtab[v][i] = valid_group_mask<MODE>(w, vtab_chunk(v, NC), i);
    std::vector<uint32_t> bits((size_t)OUT_R * WW), staged;
    const uint32_t kbias = filter_kbias((uint32_t)t);
    unsigned long long total = 0;
    uint32_t gc = 0, tag = 1;
    for (int strip = 0; strip < S; strip++) {
        staged.clear();
        for (int c = 0; c < NC; c++, gc++) {
            const ChunkGeo g = make_geo<MODE>(w, h, strip, c, SR);
            const int ty0 = g.ys0 - 3;
            if (g.xt0 % 16 != 0) return -16;  // TMA: innermost box start must be 16-byte aligned
            for (int r = 0; r < TR; r++)      // what the TMA tiled load delivers: zero fill outside the image
                for (int j = 0; j < kTileW; j++) {
                    const int y = ty0 + r, x = g.xt0 + j;
                    tile[r * kTileW + j] = (y >= 0 && y < h && x >= 0 && x < w) ? img[(size_t)y * pitch + x] : 0;
                }
            const uint32_t *vt = vtab[vtab_variant(c, NC)];
            uint32_t qcount = 0, nent[kFilterWarps];
            for (int warp = 0; warp < kFilterWarps; warp++)
                nent[warp] = phase_a_stage1<MODE, SR, kFilterWarps>(warp, -1, tile, ents.data() + warp * kWarpQueueCap, g, kbias);
            for (int ftid = 0; ftid < kFilterThreads; ftid++)
                phase_a_stage2<MODE, SR, kFilterWarps>(ftid, kFilterThreads, tile, ents.data(), nent, vt, queue.data(), &qcount, kbias);
            const uint32_t qn = qcount;
            if (tag == 1u && gc != 0u) std::fill(plane.begin(), plane.end(), (uint16_t)0);
            if (qn <= (uint32_t)kQueueCap) {
                uint32_t kn = 0;
                for (int tid = 0; tid < kTestThreads; tid++)
                    phase_b<MODE, SR>(tid, -1, kTestThreads, qn, tile, queue.data(), klist.data(), &kn, plane.data(), t, n, tag);
                if (kn > qn) return -19;
                std::vector<uint32_t> run(kn + 1);
                uint32_t scount = 0;
                bool dropped = false;
                for (int tid = 0; tid < kTestThreads; tid++)
                    dropped |= emit_list<MODE, SR>(tid, kTestThreads, kn, klist.data(), plane.data(), &scount, 0ull, kn, run.data(), g);
                if (scount > kn || dropped) return -20;
                staged.insert(staged.end(), run.begin(), run.begin() + scount);
            } else {
                if (fallbacks) fallbacks[0]++;
                for (int twarp = 0; twarp < kTestWarps; twarp++)
                    phase_b_dense<MODE, SR>(twarp, -1, kTestWarps, tile, plane.data(), g, c, t, n, tag);
                if (fallbacks) fallbacks[1]++;
                uint32_t counter = 0;
                for (int tid = 0; tid < kTestThreads; tid++)
                    nms_dense<MODE, SR>(tid, kTestThreads, 0, plane.data(), &counter, 0ull, 0ull, nullptr, g, tag);
                std::vector<uint32_t> run(counter + 1);
                const uint32_t kn = counter;
                counter = 0;
                bool dropped = false;
                for (int tid = 0; tid < kTestThreads; tid++)
                    dropped |= nms_dense<MODE, SR>(tid, kTestThreads, 1, plane.data(), &counter, 0ull, kn, run.data(), g, tag);
                if (counter != kn || dropped) return -21;
                staged.insert(staged.end(), run.begin(), run.begin() + kn);
            }
            tag = tag == (uint32_t)kTagPeriod ? 1u : tag + 1u;
        }
        // the gather kernel's part: the strip's runs -> bit plane -> row-major points
        std::fill(bits.begin(), bits.end(), 0u);
        for (uint32_t e : staged) {
            const uint32_t x = e & 0xffffu, row = e >> 16;
            if ((int)row >= OUT_R || (int)x >= w) return -22;
            if (bits[row * WW + (x >> 5)] & (1u << (x & 31u))) return -23;  // a keypoint staged twice
            bits[row * WW + (x >> 5)] |= 1u << (x & 31u);
        }
        const int y0 = first_out_row(MODE) + strip * OUT_R;
        for (int i = 0; i < OUT_R * WW; i++) {
            const uint32_t m = bits[i];
            const int row = i / WW, col = i - row * WW;
            emit_word(m, (uint32_t)col * 32u, (uint32_t)(y0 + row), total, cap, out);
            total += (unsigned long long)__builtin_popcount(m);
        }
    }
    return (int64_t)total;
}

uint64_t splitmix(uint64_t &s) {
    s += 0x9E3779B97F4A7C15ULL;
    uint64_t z = s;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

}  // namespace

extern "C" {

int64_t fdf_emulate_detect(const uint8_t *img, uint32_t w, uint32_t h, uint32_t pitch, uint8_t t, uint8_t n,
                           uint8_t nms, int sr, fdf_oracle_point *out, size_t cap, int *fallbacks) {
    uint2 *o = reinterpret_cast<uint2 *>(out);
    fdf::check_failure_line() = 0;
    if (fallbacks) fallbacks[0] = fallbacks[1] = 0;  // chunks that took the dense path: [0] every-pixel test, [1] plane scan
#define CASE(M, S)                                                                                     \
    if (nms == M && sr == S) {                                                                         \
        const int64_t k = emulate<M, S>(img, (int)w, (int)h, (int)pitch, t, n, o, cap, fallbacks);     \
        return fdf::check_failure_line() != 0 ? -1000 - fdf::check_failure_line() : k; /* an index check failed */ \
    }
    CASE(0, 32) CASE(0, 48) CASE(0, 64) CASE(1, 32) CASE(1, 48) CASE(1, 64) CASE(2, 32) CASE(2, 48) CASE(2, 64)
#undef CASE
    return -1;
}

// Random (centre, ring, t, n) vectors: device-side arithmetic of fdf_core.cuh against the oracle.
// Returns the number of disagreements (0 expected).  `style` biases the vectors towards arcs.
int64_t fdf_core_check(uint64_t iterations, uint64_t seed) {
    int64_t bad = 0;
    uint64_t s = seed;
    for (uint64_t it = 0; it < iterations; it++) {
        const uint64_t r0 = splitmix(s), r1 = splitmix(s), r2 = splitmix(s);
        const int style = (int)(r0 & 3);
        const int c = (int)((r0 >> 8) & 0xff);
        int t = (int)((r0 >> 16) & 0xff);
        if (style != 0) t &= 0x3f;
        const int n = 9 + (int)((r0 >> 24) % 8);
        uint8_t ring8[16];
        int ring[16];
        // style 0: uniform random; 1..3: an arc of random length / sign with jitter plus background
        const int arc_len = (int)((r0 >> 32) % 17), arc_start = (int)((r0 >> 40) & 15);
        const int sign = ((r0 >> 44) & 1) ? 1 : -1;
        for (int i = 0; i < 16; i++) {
            int v;
            const int rnd = (int)((i < 8 ? (r1 >> (8 * i)) : (r2 >> (8 * (i - 8)))) & 0xff);
            if (style == 0) {
                v = rnd;
            } else {
                const bool in_arc = ((i - arc_start) & 15) < arc_len;
                const int jitter = (rnd & 7) - 3;
                v = in_arc ? c + sign * (t + 1 + (rnd >> 5) * (style == 3 ? 9 : 1)) + (style == 2 ? jitter : 0)
                           : c + jitter * (style == 3 ? 5 : 1);
                v = v < 0 ? 0 : (v > 255 ? 255 : v);
            }
            ring8[i] = (uint8_t)v;
            ring[i] = v;
        }
        // oracle
        uint8_t neg[16], pos[16];
        for (int i = 0; i < 16; i++) {
            const int d = c - ring[i];
            neg[i] = d < 0 && -d > t;
            pos[i] = d > 0 && d > t;
        }
        const bool kp_bright = fdf_oracle_consecutive(neg, 16, n), kp_dark = fdf_oracle_consecutive(pos, 16, n);
        // device arithmetic (two ring pixels per word: ring[i] | ring[i + 8] << 16)
        Ring2 rp;
        for (int i = 0; i < 8; i++) rp.p[i] = (uint32_t)ring[i] | ((uint32_t)ring[i + 8] << 16);
        const RingMasks rm = ring_masks(c, rp, t);
        uint32_t mb = 0, md = 0;
        for (int i = 0; i < 16; i++) {
            mb |= (uint32_t)neg[i] << i;
            md |= (uint32_t)pos[i] << i;
        }
        if ((rm.bright & 0xffffu) != mb || (rm.dark & 0xffffu) != md) bad++;
        // the form the kernel uses: one dual word per ring pixel, test and MaxThreshold score from the best window
        RingDual rd;
        for (int i = 0; i < 16; i++) rd.w[i] = dual_word((uint32_t)ring[i], dual_bias(c));
        const uint32_t best = best_of_lanes(best_window<0>(rd, n));
        if (n == 9 && best_window<9>(rd, n) != best_window<0>(rd, n)) bad++;  // the kernel's n = 9 instantiation
        if ((best > (uint32_t)(256 + t)) != (kp_bright || kp_dark)) bad++;
        if ((kp_bright || kp_dark) &&
            best - 256u != fdf_oracle_score_max_threshold_px((uint8_t)c, ring8, (uint8_t)n))
            bad++;
        if (score_sum_abs_dual(rd, t) != fdf_oracle_score_sum_abs_px((uint8_t)c, ring8, (uint8_t)t)) bad++;
        const bool ab = has_arc(rm.bright & 0xffffu, n), ad = has_arc(rm.dark & 0xffffu, n);
        if (ab != kp_bright || ad != kp_dark) bad++;
        if (score_sum_abs(c, rp, t) != fdf_oracle_score_sum_abs_px((uint8_t)c, ring8, (uint8_t)t)) bad++;
        if (ab || ad) {
            if (score_max_threshold(c, rp, n, ab) != fdf_oracle_score_max_threshold_px((uint8_t)c, ring8, (uint8_t)n))
                bad++;
            // the filter must pass every keypoint: put the centre in each byte lane in turn
            // (a 16-pixel row group of random bytes with the keypoint's N / S / E / W planted around column px)
            for (int px = 0; px < 16; px += 5) {
                uint8_t rows[3][24];  // north, centre, south; columns -4 .. 19 of the group
                for (int rr = 0; rr < 3; rr++)
                    for (int i = 0; i < 24; i += 8) {
                        const uint64_t v = splitmix(s);
                        memcpy(&rows[rr][i], &v, 8);
                    }
                rows[1][4 + px] = (uint8_t)c;
                rows[0][4 + px] = ring8[0];
                rows[2][4 + px] = ring8[8];
                rows[1][4 + px + 3] = ring8[4];
                rows[1][4 + px - 3] = ring8[12];
                Px16 pc, pn, ps;
                memcpy(pn.w, &rows[0][4], 16);
                memcpy(pc.w, &rows[1][4], 16);
                memcpy(ps.w, &rows[2][4], 16);
                uint32_t cl, cr;
                memcpy(&cl, &rows[1][0], 4);
                memcpy(&cr, &rows[1][20], 4);
                const uint32_t all = 0xf0f0f0f0u;
                const uint32_t kb = filter_kbias((uint32_t)t);
                if (vertical_any(pc, pn, ps, kb) == 0u) bad++;
                const uint32_t m = candidate_mask16(pc, pn, ps, cl, cr, all, kb);
                bool found = false;
                for (int p = 0; p < 32; p++)
                    if (((m >> p) & 1u) && mask_bit_to_px(p) == px) found = true;
                if (!found) bad++;
            }
        }
    }
    return bad;
}

}  // extern "C"
