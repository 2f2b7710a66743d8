// agar_host_expand.h - the host side of agar_step_host's compact copy: plain C++ (no CUDA), included by agar_b200.cu and by
// the CPU test harness (tests/host_expand_harness.cpp), so that the grid rebuild and the in-place update are covered by
// the CPU suite against the oracle's dense grids.
#ifndef AGAR_HOST_EXPAND_H
#define AGAR_HOST_EXPAND_H

#include <immintrin.h>

#include <cstdint>
#include <cstring>

#include "../../include/agar_b200.h"

namespace agar {

struct ObsEntry { uint32_t idx; float val; };  // one entry of an agent's image that is neither +0.0 nor -1.0

// what the host needs to know about the grids of a handle (DevParams without the device-only fields)
struct GridGeo {
    int G, C, K, obs_flags;  // grid side, channels, players per arena, AGAR_OBS_* flags
    float arena;             // arena_size
    const float* oob_off;    // [G] (float)((double)(i - G/2) * (view/G)), features/extractors.py:87-91
};

// ---- host side of the compact copy: rebuild one agent's dense grid from its centroid and its occupied bins ----
// non-temporal copy of one channel image (no read-for-ownership of the destination lines); n % 16 == 0, src 64-byte aligned
__attribute__((target("avx512f"))) inline void stream_out_avx512(float* dst, const float* src, int n) {
    for (int i = 0; i < n; i += 16) _mm512_stream_ps(dst + i, _mm512_load_ps(src + i));
}
__attribute__((target("avx2"))) inline void stream_out_avx2(float* dst, const float* src, int n) {
    for (int i = 0; i < n; i += 8) _mm256_stream_ps(dst + i, _mm256_load_ps(src + i));
}
inline void stream_out_sse2(float* dst, const float* src, int n) {
    for (int i = 0; i < n; i += 4) _mm_stream_ps(dst + i, _mm_load_ps(src + i));
}
inline void stream_out(float* dst, const float* src, int n) {  // n = G*G, a multiple of 16
    static const int isa = __builtin_cpu_supports("avx512f") ? 2 : (__builtin_cpu_supports("avx2") ? 1 : 0);
    const uintptr_t d = reinterpret_cast<uintptr_t>(dst);
    if (isa == 2 && (d & 63u) == 0) stream_out_avx512(dst, src, n);
    else if (isa >= 1 && (d & 31u) == 0) stream_out_avx2(dst, src, n);
    else if ((d & 15u) == 0) stream_out_sse2(dst, src, n);
    else std::memcpy(dst, src, (size_t)n * sizeof(float));
}

// Same rule as the raster (features/extractors.py:84-96): grid index i is out of bounds when its world corner is.
// Each channel is assembled in a cache-resident staging image (base rows, then the channel's occupied bins) and
// leaves with non-temporal stores: the destination lines are written once and never read.
inline void expand_agent(const GridGeo& p, float* dst, const ObsEntry* ent, int count, const float* meta) {
    const int G = p.G, GG = G * G, C = p.C, flags = p.obs_flags;
    const float cx = meta[0], cy = meta[1], arena = p.arena;
    const bool alive = meta[2] != 0.0f;
    int sentmask = 0;
    {
        int ch = 0;
        if (flags & AGAR_OBS_PELLETS) { sentmask |= 1 << ch; ++ch; }
        if (flags & AGAR_OBS_CELLS) { sentmask |= 1 << ch; ++ch; }
        if (flags & AGAR_OBS_OTHERS) { if (p.K > 1) sentmask |= 1 << ch; ++ch; }  // no sentinel pass without others
        if (flags & AGAR_OBS_VIRUSES) { sentmask |= 1 << ch; ++ch; }
        if (flags & AGAR_OBS_FOODS) { sentmask |= 1 << ch; ++ch; }
    }
    const int smask = alive ? sentmask : 0;  // an agent without cells observes zeros
    alignas(64) float stage[128 * 128];
    float neg_row[128], col_row[128];
    bool oob_x[128];
    bool any_oob = false;
    for (int i = 0; i < G; ++i) {
        neg_row[i] = -1.0f;
        const float wy = p.oob_off[i] + cy, wx = p.oob_off[i] + cx;
        col_row[i] = !(0.0f <= wy && wy < arena) ? -1.0f : 0.0f;
        oob_x[i] = !(0.0f <= wx && wx < arena);
        any_oob = any_oob || oob_x[i] || col_row[i] != 0.0f;
    }
    for (int ch = 0; ch < C; ++ch) {
        std::memset(stage, 0, (size_t)GG * sizeof(float));
        if (((smask >> ch) & 1) && any_oob)
            for (int gx = 0; gx < G; ++gx) std::memcpy(stage + (size_t)gx * G, oob_x[gx] ? neg_row : col_row, (size_t)G * sizeof(float));
        const uint32_t lo = (uint32_t)ch * (uint32_t)GG;
        for (int e = 0; e < count; ++e) {
            const uint32_t rel = ent[e].idx - lo;
            if (rel < (uint32_t)GG) stage[rel] = ent[e].val;
        }
        stream_out(dst + (size_t)ch * GG, stage, GG);
    }
    _mm_sfence();
}

// ---- persistent host buffer: bring one agent's grid from the previous step's image to this step's in place ----
inline int sentinel_mask(const GridGeo& p) {  // channels that carry the -1 out-of-bounds marks (same rule as expand_agent)
    const int flags = p.obs_flags;
    int m = 0, ch = 0;
    if (flags & AGAR_OBS_PELLETS) { m |= 1 << ch; ++ch; }
    if (flags & AGAR_OBS_CELLS) { m |= 1 << ch; ++ch; }
    if (flags & AGAR_OBS_OTHERS) { if (p.K > 1) m |= 1 << ch; ++ch; }
    if (flags & AGAR_OBS_VIRUSES) { m |= 1 << ch; ++ch; }
    if (flags & AGAR_OBS_FOODS) { m |= 1 << ch; ++ch; }
    return m;
}
// What changes between the image described by the previous step's (list, centroid) and this step's.  The world
// coordinate of grid index i, oob_off[i] + c, does not decrease with i, so the out-of-bounds indices are a prefix
// (coordinate < 0) and a suffix (coordinate >= arena): index i is in bounds iff lo <= i < hi.
struct DeltaPlan {
    int xlo, xhi, ylo, yhi;  // this step's in-bounds rows (x) and columns (y)
    int rr[2][2];            // two ranges of rows whose state flipped: rewritten whole
    uint8_t chg[128];        // columns whose state flipped: patched in every unchanged in-bounds row
    int nchg, smask;
    bool rows, any, full;    // some row flipped | some index out of bounds now | alive flag flipped: full rewrite
};
inline void inbounds_range(const GridGeo& p, float c, int& lo, int& hi) {
    const int G = p.G;
    const float arena = p.arena;
    int i = 0;
    while (i < G && !(0.0f <= p.oob_off[i] + c)) ++i;   // (the same float expression as the raster's rule)
    lo = i;
    int j = G;
    while (j > lo && !(p.oob_off[j - 1] + c < arena)) --j;
    hi = j;
}
inline void delta_plan(const GridGeo& p, const float* pmeta, const float* nmeta, int sentmask, DeltaPlan& d) {
    const int G = p.G;
    const int sm0 = pmeta[2] != 0.0f ? sentmask : 0, sm1 = nmeta[2] != 0.0f ? sentmask : 0;  // an agent without cells observes zeros
    d.full = sm0 != sm1; d.smask = sm1; d.nchg = 0; d.rows = false; d.any = false;
    if (d.full || !sm1) return;
    int xlo0, xhi0, ylo0, yhi0;
    inbounds_range(p, pmeta[0], xlo0, xhi0); inbounds_range(p, pmeta[1], ylo0, yhi0);
    inbounds_range(p, nmeta[0], d.xlo, d.xhi); inbounds_range(p, nmeta[1], d.ylo, d.yhi);
    d.any = d.xlo > 0 || d.xhi < G || d.ylo > 0 || d.yhi < G;
    // index i flipped iff it lies between the old and the new bound (lower bounds, upper bounds); when a range is empty
    // on one side (lo == hi: everything out of bounds) the two intervals may overlap, which only repeats work
    d.rr[0][0] = xlo0 < d.xlo ? xlo0 : d.xlo; d.rr[0][1] = xlo0 < d.xlo ? d.xlo : xlo0;
    d.rr[1][0] = xhi0 < d.xhi ? xhi0 : d.xhi; d.rr[1][1] = xhi0 < d.xhi ? d.xhi : xhi0;
    d.rows = d.rr[0][0] < d.rr[0][1] || d.rr[1][0] < d.rr[1][1];
    for (int i = (ylo0 < d.ylo ? ylo0 : d.ylo); i < (ylo0 < d.ylo ? d.ylo : ylo0); ++i) d.chg[d.nchg++] = (uint8_t)i;
    for (int i = (yhi0 < d.yhi ? yhi0 : d.yhi); i < (yhi0 < d.yhi ? d.yhi : yhi0); ++i)
        if (d.nchg < 128 && !(d.nchg && i <= d.chg[d.nchg - 1])) d.chg[d.nchg++] = (uint8_t)i;
}
// Prefetch (for writing) the lines delta_apply will touch: issued for a worker's next agent while it updates the
// current one, so that the read-for-ownership misses of the scattered stores overlap.
inline void delta_prefetch(const GridGeo& p, const float* dst, const DeltaPlan& d, const ObsEntry* pe, int pc, const ObsEntry* ne, int nc) {
    if (d.full) return;
    for (int e = 0; e < pc; ++e) __builtin_prefetch(dst + pe[e].idx, 1, 3);
    for (int e = 0; e < nc; ++e) __builtin_prefetch(dst + ne[e].idx, 1, 3);
    if (!d.smask || !(d.rows || d.nchg)) return;
    const int G = p.G, GG = G * G, C = p.C;
    const int c0 = d.nchg ? d.chg[0] : 0, c1 = d.nchg ? d.chg[d.nchg - 1] : 0;
    for (int ch = 0; ch < C; ++ch) {
        if (!((d.smask >> ch) & 1)) continue;
        const float* img = dst + (size_t)ch * GG;
        for (int r = 0; r < 2; ++r)
            for (int gx = d.rr[r][0]; gx < d.rr[r][1]; ++gx)
                for (int i = 0; i < G; i += 16) __builtin_prefetch(img + (size_t)gx * G + i, 1, 3);
        if (d.nchg)
            for (int gx = d.xlo; gx < d.xhi; ++gx) {
                __builtin_prefetch(img + (size_t)gx * G + c0, 1, 3);
                if ((c1 >> 4) != (c0 >> 4)) __builtin_prefetch(img + (size_t)gx * G + c1, 1, 3);
            }
    }
}
// dst holds the image described by the previous lists; afterwards it holds this step's: the rows whose out-of-bounds
// state flipped are rewritten, the columns whose state flipped are patched row by row, the bins occupied before are
// returned to the new base value and the bins occupied now are written.  Plain (cached) stores: only the lines that
// change are touched.
inline void delta_apply(const GridGeo& p, float* dst, const DeltaPlan& d, const ObsEntry* pe, int pc, const ObsEntry* ne, int nc,
                 const float* nmeta) {
    if (d.full) { expand_agent(p, dst, ne, nc, nmeta); return; }  // the agent died or respawned: rare
    const int G = p.G, GG = G * G, C = p.C, smask = d.smask;
    if (smask && (d.rows || d.nchg)) {
        float neg_row[128], col_row[128];
        for (int i = 0; i < G; ++i) { neg_row[i] = -1.0f; col_row[i] = (i < d.ylo || i >= d.yhi) ? -1.0f : 0.0f; }
        for (int ch = 0; ch < C; ++ch) {
            if (!((smask >> ch) & 1)) continue;
            float* img = dst + (size_t)ch * GG;
            if (d.nchg)  // (rows that flipped are rewritten whole below; patching them first is harmless)
                for (int gx = d.xlo; gx < d.xhi; ++gx) {
                    float* row = img + (size_t)gx * G;
                    for (int q = 0; q < d.nchg; ++q) row[d.chg[q]] = col_row[d.chg[q]];
                }
            for (int r = 0; r < 2; ++r)
                for (int gx = d.rr[r][0]; gx < d.rr[r][1]; ++gx)
                    std::memcpy(img + (size_t)gx * G, (gx < d.xlo || gx >= d.xhi) ? neg_row : col_row, (size_t)G * sizeof(float));
        }
    }
    if (!(smask && d.any)) {  // no out-of-bounds mark anywhere in this step's image: the base is zero
        for (int e = 0; e < pc; ++e) dst[pe[e].idx] = 0.0f;
    } else {  // idx -> (channel, gx, gy) by multiply-shift (exact for idx < 2^18, divisors <= 2^14)
        const uint64_t mgg = ((uint64_t)1 << 32) / (uint64_t)GG + 1, mg = ((uint64_t)1 << 32) / (uint64_t)G + 1;
        for (int e = 0; e < pc; ++e) {
            const uint32_t idx = pe[e].idx, ch = (uint32_t)(((uint64_t)idx * mgg) >> 32), rel = idx - ch * (uint32_t)GG;
            const int gx = (int)(((uint64_t)rel * mg) >> 32), gy = (int)rel - gx * G;
            const bool out = gx < d.xlo || gx >= d.xhi || gy < d.ylo || gy >= d.yhi;
            dst[idx] = (((smask >> ch) & 1) && out) ? -1.0f : 0.0f;
        }
    }
    for (int e = 0; e < nc; ++e) dst[ne[e].idx] = ne[e].val;
}

}  // namespace agar
#endif  // AGAR_HOST_EXPAND_H
