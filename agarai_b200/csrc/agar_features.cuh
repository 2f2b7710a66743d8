// agar_features.cuh — the other two observation families of features/extractors.py on the device:
//
//   screen_gray_kernel   ScreenFeatureExtractor.extract (features/extractors.py:152-164): mean over the
//                        trailing colour axis of uint8 frames.  HBM-bound: 3 B read, 4 or 8 B written per pixel.
//   ram_features_kernel  FeatureExtractor.extract (features/extractors.py:99-149 with the helpers :171-212):
//                        nearest-k pellets / viruses / foods, the agent's cells, the closest players' cells,
//                        as one fixed-length vector per agent, straight from the arena records.
//
// Both are bit-exact restatements in float32/float64 single IEEE operations (compiled with -fmad=false); the
// sequential definitions are oracle_screen_gray / oracle_ram_extract in oracle/agar_oracle.c.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "agar_host_expand.h"
#include "agar_kernels.cuh"

namespace agar {

// ------------------------------------------------------------------------------ screen -> gray
// numpy: observation.mean(axis=-1) on uint8 adds the three bytes in float64 and divides by 3.0; the sum of
// three bytes is exact in any order, so out = (double)(r+g+b) / 3.0 is the same float64 (float32 output is that
// value rounded once, i.e. what `.astype(np.float32)` of the reference's result gives).
// A CTA converts a tile of 4096 pixels: the 12 KB of bytes are staged in shared memory with coalesced 128-bit loads,
// then every thread converts adjacent pixels (four for float32, two for float64 output) per pass so that each store
// instruction of a warp writes 512 contiguous bytes.  The quotient comes from a 766-entry table of (double)s / 3.0 built
// per CTA (the same correctly rounded IEEE division, done 3 times per thread instead of once per pixel).
#define GRAY_TILE 4096
template <typename OutT>
__global__ void __launch_bounds__(256) screen_gray_kernel(const uint8_t* __restrict__ rgb, long long n_pixels,
                                                          OutT* __restrict__ out) {
    __shared__ __align__(16) uint8_t sb[GRAY_TILE * 3];
    __shared__ OutT tab[768];  // (float entries = the float64 quotient rounded once, as `.astype(np.float32)` gives)
    const int tid = threadIdx.x;
    const long long p0 = (long long)blockIdx.x * GRAY_TILE;
    for (int i = tid; i < 766; i += 256) tab[i] = (OutT)((double)i / 3.0);
    if (p0 + GRAY_TILE <= n_pixels) {
        const uint4* src = reinterpret_cast<const uint4*>(rgb + p0 * 3);  // 16 B aligned when rgb is (a tile is 12,288 B)
        uint4* dst = reinterpret_cast<uint4*>(sb);
#pragma unroll
        for (int k = 0; k < 3; ++k) dst[k * 256 + tid] = __ldcs(src + k * 256 + tid);
        __syncthreads();
        if (sizeof(OutT) == 4) {
            const uint32_t* w = reinterpret_cast<const uint32_t*>(sb);
            float4* o = reinterpret_cast<float4*>(out + p0);
#pragma unroll
            for (int k = 0; k < GRAY_TILE / 4 / 256; ++k) {
                const int q = k * 256 + tid;  // pixels 4q .. 4q+3 = bytes 12q .. 12q+11 (three aligned words; conflict-free)
                const uint32_t a = w[3 * q], b = w[3 * q + 1], c = w[3 * q + 2];
                const int s0 = (a & 255) + ((a >> 8) & 255) + ((a >> 16) & 255);
                const int s1 = (a >> 24) + (b & 255) + ((b >> 8) & 255);
                const int s2 = ((b >> 16) & 255) + (b >> 24) + (c & 255);
                const int s3 = ((c >> 8) & 255) + ((c >> 16) & 255) + (c >> 24);
                __stcs(o + q, make_float4((float)tab[s0], (float)tab[s1], (float)tab[s2], (float)tab[s3]));  // (tab is float here)
            }
        } else {
            const uint16_t* h = reinterpret_cast<const uint16_t*>(sb);
            double2* o = reinterpret_cast<double2*>(out + p0);
#pragma unroll
            for (int k = 0; k < GRAY_TILE / 2 / 256; ++k) {
                const int q = k * 256 + tid;  // pixels 2q, 2q+1 = bytes 6q .. 6q+5 (three aligned half-words)
                const uint32_t a = h[3 * q], b = h[3 * q + 1], c = h[3 * q + 2];
                const int s0 = (a & 255) + (a >> 8) + (b & 255);
                const int s1 = (b >> 8) + (c & 255) + (c >> 8);
                __stcs(o + q, make_double2(tab[s0], tab[s1]));
            }
        }
    } else {  // the last, partial tile
        __syncthreads();
        for (long long p = p0 + tid; p < n_pixels; p += 256)
            out[p] = tab[(int)rgb[3 * p] + (int)rgb[3 * p + 1] + (int)rgb[3 * p + 2]];
    }
}

// ------------------------------------------------------------------------------ compact observations
// agar_step_host hands dense float32 grids to a host buffer.  An agent's grid (C x G x G, 48 KB at 3 x 64 x 64) is a
// base image of zeros and -1 sentinels plus ~100 occupied bins, and the sentinels follow from the agent's centroid
// (obs_meta, written by the raster).  This kernel lists the entries of an image that are neither +0.0 nor -1.0 as
// (index, value) pairs; the host copies the lists (a few KB per agent) instead of the images and rebuilds the dense
// grids with its own threads.  One CTA per agent; the order of the pairs is arbitrary (indices are distinct).
// (ObsEntry: agar_host_expand.h)
__global__ void __launch_bounds__(256) obs_compact_kernel(const float* __restrict__ obs, int floats_per_agent, int cap,
                                                         ObsEntry* __restrict__ entries, int32_t* __restrict__ counts, int count_stride) {
    __shared__ int cnt;
    const int tid = threadIdx.x, agent = blockIdx.x;
    if (tid == 0) cnt = 0;
    __syncthreads();
    const float4* img = reinterpret_cast<const float4*>(obs + (size_t)agent * floats_per_agent);
    ObsEntry* out = entries + (size_t)agent * cap;
    auto emit = [&](uint32_t idx, float v) {
        const uint32_t bits = __float_as_uint(v);
        if (bits != 0u && bits != 0xBF800000u) {  // not +0.0, not -1.0 (bit patterns: the copy must be exact)
            const int pos = atomicAdd(&cnt, 1);
            if (pos < cap) { ObsEntry e; e.idx = idx; e.val = v; out[pos] = e; }
        }
    };
    for (int i = tid; i < floats_per_agent / 4; i += 256) {
        const float4 v = __ldcs(img + i);
        emit(4u * i, v.x); emit(4u * i + 1u, v.y); emit(4u * i + 2u, v.z); emit(4u * i + 3u, v.w);
    }
    __syncthreads();
    if (tid == 0) counts[(size_t)agent * count_stride] = cnt;  // > cap: the list overflowed and the host copies this agent's image instead
}

// ------------------------------------------------------------------------------ RAM-style features
struct RamParams {
    int n_pellet, n_virus, n_food, n_other, n_cell;  // FeatureExtractor.__init__ (features/extractors.py:101-106)
    int size;                                         // 2*(n_pellet+n_virus+n_food) + 5*(1+n_other)*n_cell (:108)
    int sort_cap;                                     // power of two >= max(P, V, F): shared-memory sort buffer
    float filler;                                     // -1000 (:110)
};

// ascending bitonic sort of n (power of two) 64-bit keys in shared memory, all threads of the CTA
__device__ __forceinline__ void bitonic_sort_u64(unsigned long long* key, int n) {
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < n; i += nt) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long a = key[i], b = key[ixj];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) { key[i] = b; key[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
}

// Ranks the valid keys of key[0..C) (~0ull = none; valid keys are unique: their low half is an index) by counting, for
// each, the keys below it - no barriers, every thread reads the same key at the same time (a shared-memory broadcast)
// - and calls row(rank, key) for those ranked below n.  C*C/blockDim comparisons per thread: used for C <= RAM_RANK_CAP.
#define RAM_RANK_CAP 320
template <class RowFn>
__device__ __forceinline__ void rank_emit(const unsigned long long* key, int C, int n, RowFn row) {
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
        const unsigned long long ki = key[i];
        if (ki == ~0ull) continue;
        int rank = 0;
#pragma unroll 4
        for (int j = 0; j < C; ++j) rank += key[j] < ki ? 1 : 0;
        if (rank < n) row(rank, ki);
    }
}

// get_entity_features(loc, entities[:, :2], n) for point entities (features/extractors.py:171-180,
// :197-204): the n entities nearest to loc by float32 sqrt(dx*dx + dy*dy), ties by index (the stable order;
// numpy's default argsort leaves ties unspecified), as positions relative to loc; zero rows pad.
// key = distance bits (non-negative floats order like their bit patterns) << 32 | index.
// Select, then rank: with many more entities than wanted, only the entities inside a radius that certainly holds n of
// them are kept (seven candidate radii, multiples of the radius a uniform density would need, are counted in the pass
// that builds the keys; the smallest sufficient one selects the candidates, which are compacted); the n nearest of
// all are among them.  The kept keys are ranked by counting (rank_emit) instead of sorted; only a set larger than
// RAM_RANK_CAP (no radius qualified: a very uneven scatter, or n close to the entity count) falls back to the bitonic
// sort of everything.  (Round 1 sorted 128-512 candidates with ~30 barrier-separated passes: 175 us for 4096 agents.)
#define RAM_CAND_CAP 512
#define RAM_NRAD 7
__device__ void ram_nearest_points(unsigned long long* key, unsigned long long* cand, int* scal, int cap, const float* ex,
                                   const float* ey, int count, bool marker, float lx, float ly, int n, float arena,
                                   float* __restrict__ out) {
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
    int m = count > 1 ? 1 << (32 - __clz(count - 1)) : 1;  // next power of two (the bitonic fallback's buffer size)
    if (m > cap) m = cap;
    const bool preselect = count > 64 && n * 2 < count;  // block-uniform
    float rad[RAM_NRAD];
    {
        const float r0 = preselect ? sqrtf((float)n * arena * arena / (3.14159265f * (float)count)) : 0.0f;
        const float mult[RAM_NRAD] = {1.25f, 1.5f, 1.8f, 2.2f, 3.0f, 4.5f, 7.0f};
#pragma unroll
        for (int q = 0; q < RAM_NRAD; ++q) rad[q] = r0 * mult[q];
    }
    if (tid < RAM_NRAD + 2) scal[tid] = 0;  // [0..NRAD): entities inside each radius, [NRAD]: compaction cursor, [NRAD+1]: valid entities
    __syncthreads();
    int nval = 0, nin[RAM_NRAD];  // this thread's valid entities, and those inside each radius
#pragma unroll
    for (int q = 0; q < RAM_NRAD; ++q) nin[q] = 0;
    // four entities per thread and pass: the coordinate segments of a record are 16-byte aligned and padded to
    // multiples of four (padding entries carry x = -1 or lie beyond `count`: masked below)
    for (int base = 0; base < m; base += 4 * nt) {
        const int i0 = base + 4 * tid;
        float xs[4] = {-1.0f, -1.0f, -1.0f, -1.0f}, ys[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        if (i0 < count) {
            const float4 x4 = *reinterpret_cast<const float4*>(ex + i0), y4 = *reinterpret_cast<const float4*>(ey + i0);
            xs[0] = x4.x; xs[1] = x4.y; xs[2] = x4.z; xs[3] = x4.w;
            ys[0] = y4.x; ys[1] = y4.y; ys[2] = y4.z; ys[3] = y4.w;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int i = i0 + e;
            unsigned long long k = ~0ull;
            float d = INFINITY;
            if (i < count && (!marker || xs[e] >= 0.0f)) {
                const float dx = xs[e] - lx, dy = ys[e] - ly;
                d = sqrtf(dx * dx + dy * dy);
                k = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)i;
                ++nval;
            }
            if (i < m) key[i] = k;
            if (preselect) {
#pragma unroll
                for (int q = 0; q < RAM_NRAD; ++q) nin[q] += d <= rad[q] ? 1 : 0;
            }
        }
    }
    {   // one reduction per warp and counter
        const int tv = __reduce_add_sync(FULL, nval);
        if (lane == 0 && tv) atomicAdd(&scal[RAM_NRAD + 1], tv);
        if (preselect) {
#pragma unroll
            for (int q = 0; q < RAM_NRAD; ++q) {
                const int tq = __reduce_add_sync(FULL, nin[q]);
                if (lane == 0 && tq) atomicAdd(&scal[q], tq);
            }
        }
    }
    __syncthreads();
    const unsigned long long* keys = key;
    int C = m, nvalid = scal[RAM_NRAD + 1];
    if (preselect) {
        int sel = -1;
#pragma unroll
        for (int q = RAM_NRAD - 1; q >= 0; --q)
            if (scal[q] >= n && scal[q] <= RAM_CAND_CAP) sel = q;  // smallest sufficient radius
        if (sel >= 0) {  // block-uniform
            const unsigned lim = __float_as_uint(rad[sel]);
            for (int base = 0; base < m; base += nt) {  // warp-uniform; one cursor update per warp
                const int i = base + tid;
                const unsigned long long k = i < m ? key[i] : ~0ull;
                const bool keep = (unsigned)(k >> 32) <= lim;   // (~0ull: distance bits above every radius)
                const unsigned kb = __ballot_sync(FULL, keep);
                int pos = 0;
                if (lane == 0 && kb) pos = atomicAdd(&scal[RAM_NRAD], __popc(kb));
                pos = __shfl_sync(FULL, pos, 0);
                if (keep) cand[pos + __popc(kb & ((1u << lane) - 1u))] = k;  // any order: ranks are computed
            }
            __syncthreads();
            keys = cand; C = scal[sel]; nvalid = C;
        }
    }
    auto row = [&](int r, unsigned long long k) {
        const int i = (int)(unsigned)k;
        out[2 * r] = ex[i] - lx; out[2 * r + 1] = ey[i] - ly;
    };
    if (C <= RAM_RANK_CAP) {
        rank_emit(keys, C, n, row);
    } else {  // rare: sort everything (needs a power-of-two buffer: the candidates are padded in place)
        int ms = 1;
        while (ms < C) ms <<= 1;
        unsigned long long* buf = keys == cand ? cand : key;
        for (int i = C + tid; i < ms; i += nt) buf[i] = ~0ull;
        __syncthreads();
        bitonic_sort_u64(buf, ms);
        for (int r = tid; r < n && r < nvalid; r += nt) row(r, buf[r]);
    }
    for (int r = nvalid + tid; r < n; r += nt) { out[2 * r] = 0.0f; out[2 * r + 1] = 0.0f; }  // zero rows pad
    __syncthreads();
}

// One warp: largest_cells(player, n) then get_entity_features(loc, cells, n, relative) for one player's 16
// cell slots (lane = slot; lanes >= 16 idle).  features/extractors.py:182-184 sorts ASCENDING by mass and keeps
// the first n (so it keeps the n smallest: reference quirk, reproduced), :197-200 then orders those by distance
// to loc.  Both sorts are stable here (ties keep the earlier order).  Rows are [x, y, ix, iy, mass] (the
// columns 2-3 the reference never interprets carry the cell's launch impulse).  Zero rows pad.
__device__ void ram_player_cells(const float* cx, const float* cy, const float* cix, const float* ciy, const float* cm,
                                 float lx, float ly, bool relative, int n, float* __restrict__ out, int lane) {
    const bool in = lane < MC;
    const float m = in ? cm[lane] : 0.0f;
    const float x = in ? cx[lane] : 0.0f, y = in ? cy[lane] : 0.0f;
    const bool alive = m > 0.0f;
    const unsigned aliveb = __ballot_sync(FULL, alive);
    if (__popc(aliveb) <= 1) {  // the usual player: one cell (or none), nothing to rank
        if (alive && n > 0) {
            out[0] = relative ? x - lx : x; out[1] = relative ? y - ly : y;
            out[2] = cix[lane]; out[3] = ciy[lane]; out[4] = m;
        }
        for (int r = (aliveb && n > 0 ? 5 : 0) + lane; r < n * 5; r += 32) out[r] = 0.0f;
        return;
    }
    // rank by (mass, slot) among alive cells
    int rank_m = 0;
    for (int j = 0; j < MC; ++j) {
        const float mj = __shfl_sync(FULL, m, j);
        if (mj > 0.0f && (mj < m || (mj == m && j < lane))) ++rank_m;
    }
    const bool sel = alive && rank_m < n;
    const float dx = x - lx, dy = y - ly;
    const float d = sqrtf(dx * dx + dy * dy);
    // rank by (distance, mass rank) among the selected
    int rank_d = 0;
    for (int j = 0; j < MC; ++j) {
        const float dj = __shfl_sync(FULL, d, j);
        const int rj = __shfl_sync(FULL, rank_m, j);
        const bool sj = __shfl_sync(FULL, sel ? 1 : 0, j) != 0;
        if (sj && (dj < d || (dj == d && rj < rank_m))) ++rank_d;
    }
    const int nsel = __popc(__ballot_sync(FULL, sel));
    if (sel) {
        float* o = out + 5 * rank_d;
        o[0] = relative ? dx : x; o[1] = relative ? dy : y;
        o[2] = cix[lane]; o[3] = ciy[lane]; o[4] = m;
    }
    for (int r = nsel * 5 + lane; r < n * 5; r += 32) out[r] = 0.0f;
}

// grid = (arenas * agents); block = 128; dynamic smem = (sort_cap + 64 + RAM_CAND_CAP) * 8 bytes.
#ifndef RAM_MIN_BLOCKS
#define RAM_MIN_BLOCKS 10  // 51 registers: 10 CTAs per SM = 2.8 rounds for 4096 agents (9 CTAs: a fourth, nearly empty round)
#endif
__global__ void __launch_bounds__(128, RAM_MIN_BLOCKS) ram_features_kernel(const __grid_constant__ DevParams p, const __grid_constant__ RamParams rp,
                                                           const float* __restrict__ state, float* __restrict__ out_all) {
    extern __shared__ __align__(16) unsigned long long ram_smem[];
    unsigned long long* key = ram_smem;              // [sort_cap]
    unsigned long long* pkey = ram_smem + rp.sort_cap;  // [64] other players by distance
    unsigned long long* cand = pkey + 64;            // [RAM_CAND_CAP] preselected candidates
    __shared__ int s_scal[12];
    __shared__ float s_loc[2];
    __shared__ int s_has;
    const AgarLayout& L = p.L;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const int A = p.c.num_agents, K = p.K;
    const int arena = blockIdx.x / A, a = blockIdx.x - arena * A;
    const float* rec = state + (size_t)arena * L.words;
    float* out = out_all + (size_t)blockIdx.x * rp.size;
    const float* cx = rec + L.cell_x; const float* cy = rec + L.cell_y; const float* cix = rec + L.cell_ix;
    const float* ciy = rec + L.cell_iy; const float* cm = rec + L.cell_m;

    // loc = position(agent) (features/extractors.py:207-212)
    if (warp == 0) {
        const int g = a * MC + (lane & 15);
        const bool v = lane < MC;
        float lx = 0.0f, ly = 0.0f;
        const bool has = hw_centroid(0, lane & 15, v ? cx[g] : 0.0f, v ? cy[g] : 0.0f, v ? cm[g] : 0.0f, lx, ly);
        if (lane == 0) { s_loc[0] = lx; s_loc[1] = ly; s_has = has ? 1 : 0; }
    }
    __syncthreads();
    if (!s_has) {  // the reference returns None; the fixed-shape buffer gets zeros
        for (int i = tid; i < rp.size; i += nt) out[i] = 0.0f;
        return;
    }
    const float lx = s_loc[0], ly = s_loc[1];
    int off = 0;
    ram_nearest_points(key, cand, s_scal, rp.sort_cap, rec + L.pellet_x, rec + L.pellet_y, p.c.num_pellets, true, lx, ly, rp.n_pellet, p.c.arena_size, out + off);
    off += 2 * rp.n_pellet;
    ram_nearest_points(key, cand, s_scal, rp.sort_cap, rec + L.virus_x, rec + L.virus_y, p.c.num_viruses, false, lx, ly, rp.n_virus, p.c.arena_size, out + off);
    off += 2 * rp.n_virus;
    ram_nearest_points(key, cand, s_scal, rp.sort_cap, rec + L.food_x, rec + L.food_y, p.c.max_foods, true, lx, ly, rp.n_food, p.c.arena_size, out + off);
    off += 2 * rp.n_food;
    // the agent's own cells: absolute coordinates (relative=False, :131)
    if (warp == 0)
        ram_player_cells(cx + a * MC, cy + a * MC, cix + a * MC, ciy + a * MC, cm + a * MC, lx, ly, false, rp.n_cell, out + off, lane);
    off += 5 * rp.n_cell;
    // closest_players (:186-193): the other players that have cells, by distance between centroids
    // (float32 sqrt(dx*dx + dy*dy)), ties by player index
    for (int k0 = 2 * warp; k0 < 64; k0 += 2 * (nt >> 5)) {  // two players per warp pass, a half-warp each
        const int hb = lane & 16, sl = lane & 15, k = k0 + (hb >> 4);
        const bool v = k < K && k != a;
        const int g = k * MC + sl;
        const float m = v ? cm[g] : 0.0f, x = v ? cx[g] : 0.0f, y = v ? cy[g] : 0.0f;
        const int ncell = __popc((__ballot_sync(FULL, m > 0.0f) >> hb) & 0xFFFFu);
        float qx = 0.0f, qy = 0.0f;
        if (__any_sync(FULL, ncell >= 2)) {  // warp-uniform: numpy's pairwise order for several cells
            hw_centroid(hb, sl, x, y, m, qx, qy);
        } else if (m > 0.0f) {               // np.average over one cell: (x*m)/m, as the general path gives
            qx = (x * m) / m; qy = (y * m) / m;
        }
        // the lane holding a cell (any cell: the centroid is the same in all lanes of the general path) writes the key
        const unsigned alive = (__ballot_sync(FULL, m > 0.0f) >> hb) & 0xFFFFu;
        if (sl == (alive ? __ffs(alive) - 1 : 0)) {
            unsigned long long kk = ~0ull;
            if (alive) {
                const float dx = lx - qx, dy = ly - qy;
                kk = ((unsigned long long)__float_as_uint(sqrtf(dx * dx + dy * dy)) << 32) | (unsigned)k;
            }
            pkey[k] = kk;
        }
    }
    __syncthreads();
    {   // order the (at most 64) players by counting ranks: sorted[rank] = key, then a barrier
        unsigned long long mine = ~0ull;
        int rank = 64;
        if (tid < 64) {
            mine = pkey[tid];
            rank = 0;
            for (int j = 0; j < K; ++j) rank += pkey[j] < mine ? 1 : 0;  // (entries from K on are ~0ull)
        }
        __syncthreads();
        if (tid < 64) pkey[tid] = ~0ull;
        __syncthreads();
        if (mine != ~0ull) pkey[rank] = mine;  // valid keys are unique: distinct ranks
        __syncthreads();
    }
    for (int j = warp; j < rp.n_other; j += nt >> 5) {  // one warp per selected player
        float* o = out + off + 5 * rp.n_cell * j;
        const unsigned long long kk = j < 64 ? pkey[j] : ~0ull;
        if (kk == ~0ull) {  // not enough players: filler rows (:139-141)
            for (int r = lane; r < 5 * rp.n_cell; r += 32) o[r] = rp.filler;
        } else {
            const int k = (int)(unsigned)kk;
            ram_player_cells(cx + k * MC, cy + k * MC, cix + k * MC, ciy + k * MC, cm + k * MC, lx, ly, true, rp.n_cell, o, lane);
        }
    }
}

}  // namespace agar
