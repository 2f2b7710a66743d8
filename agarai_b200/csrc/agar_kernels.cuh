// agar_kernels.cuh — device side of the B200 batched Agar.io engine (sm_100a).
//
// One CTA owns one arena for one env step: it stages the arena's whole state record in shared
// memory, runs ticks_per_step ticks there, writes back only what changed and rasterises the
// grid observations straight from shared memory (base image from registers with 128-bit stores,
// occupied bins patched in place).  The rules are
// those of DESIGN.md §3; the sequential definition is oracle/agar_oracle.c and the two must
// agree bit for bit, so every float expression here is written as single IEEE operations in
// the same order (the file is compiled with -fmad=false; sqrtf and '/' are the correctly
// rounded forms under nvcc's default -prec-sqrt/-prec-div).
//
// Reference interfaces this stands in for: env.step / env.reset (ppo/train.py:164,173,
// a2c/training.py:66,94, dqn/training.py:96,104) and GridFeatureExtractor.extract
// (features/extractors.py:39-96).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/agar_b200.h"

#define MC AGAR_MAX_CELLS
#define NONE_I 0x7fffffff
#define NONE16 0xFFFFu

// Tuning builds only (-DAGAR_PHASE_CLOCKS): thread 0 of every CTA charges the cycles since the previous mark to
// a phase id; totals are read back with agar_debug_phase_clocks().  Compiled out of the product.
#ifdef AGAR_PHASE_CLOCKS
__device__ unsigned long long g_phase_clocks[32];
__device__ unsigned int g_arena_cycles[8192];  // cycles of the last launch's CTA of each of the first 8192 arenas
__shared__ long long ph_acc[33];  // [32] = time of the previous mark
#define PHASE_DECL if (threadIdx.x == 0) { for (int i_ = 0; i_ < 32; ++i_) ph_acc[i_] = 0; ph_acc[32] = clock64(); }
#define PHASE(id) do { if (threadIdx.x == 0) { const long long t_ = clock64(); ph_acc[id] += t_ - ph_acc[32]; ph_acc[32] = t_; } } while (0)
#define PHASE_FLUSH do { if (threadIdx.x == 0) { long long tot_ = 0; for (int i_ = 0; i_ < 32; ++i_) if (ph_acc[i_]) { atomicAdd(&g_phase_clocks[i_], (unsigned long long)ph_acc[i_]); tot_ += ph_acc[i_]; } if (arena < 8192) g_arena_cycles[arena] = (unsigned int)tot_; } } while (0)
#else
#define PHASE_DECL
#define PHASE(id) do { } while (0)
#define PHASE_FLUSH do { } while (0)
#endif

namespace agar {

struct DevParams {
    AgarConfig c;
    AgarLayout L;
    uint32_t seed_lo, seed_hi;
    int C;            // observation channels
    float box;        // (float)(view/G), the divisor of features/extractors.py:77-78
    int K, KC;
    float hash_inv;   // HASH_W / arena_size: pellet bin index = (int)(coordinate * hash_inv)
    float hash_bs;    // arena_size / HASH_W
    int hash_epoch;   // persisted hashes carrying another epoch are stale (bumped by agar_set_state / agar_seed)
    int hash_words;   // words of one arena's persisted hash
    float box_inv;    // 1/box when box is a power of two (then x*box_inv == x/box exactly), else 0
    float raster_lim; // entities farther than this from the centre on either axis cannot land in the grid
    float oob_off[128];  // (float)((double)(i - G/2) * (view/G)), features/extractors.py:87-91
};

struct KArgs {
    float* state;            // [N, words] world state (library owned)
    const float* src_state;  // records to read (== state except for observe_records)
    const float* target_xy;  // [N,A,2]
    const int32_t* act;      // [N,A]
    const int32_t* act_index;  // [N,A] or null
    const float* tab_xy;       // [n_actions,2]
    const int32_t* tab_act;    // [n_actions]
    int32_t n_actions;
    float* obs;              // [N,A,C,G,G] or null
    float* obs_meta;         // [N,A,4] or null: per agent (centroid x, y, has cells, 0), for agar_step_host's compact copy
    float* reward;           // [N,A]
    uint8_t* done;           // [N,A]
    const uint8_t* reset_mask;  // [N] or null (= all) — only read when F_RESET
    int32_t* events;         // [N,cap,4]
    int32_t* ev_counts;      // [N]
    int32_t* hash;           // [N, hash_words] persisted pellet hash (null: rebuild every step)
    int32_t flags;
    int32_t n_records;
    int32_t arena0;          // first arena of this launch (agar_step_host steps the batch in chunks)
    // Costly arenas first (full-batch step launches only; all null otherwise).  Every CTA measures its own duration and
    // files its arena into one of four lists of the order written for the next launch, by the ratio to the mean CTA
    // duration of the previous launch (> 2x, > 1.25x, > 0.85x, rest); the next launch walks the lists from the costliest,
    // so that the few very costly arenas of an aged batch start in the first round instead of setting the launch's
    // tail.  An order buffer is {header[8], list[4][order_stride]}; header = {count of class 3, 2, 1, 0, cycle sum
    // (64 bit), 0, 0}.  Three buffers rotate: block 0 of a launch clears the header of the one the next launch writes.
    const int32_t* order_in; // order consumed by this launch, or null (= identity)
    int32_t* order_out;      // order written by this launch
    int32_t* order_clear;    // order whose header this launch clears (written by the next launch)
    int32_t order_stride;    // capacity of one list (>= N)
};

enum { F_STEP = 1, F_RESET = 2, F_OBS = 4, F_READONLY = 8 };

// ------------------------------------------------------------------------- config views
// The kernels are templates over a "config view".  RtCfg reads every size and record offset from the kernel
// parameters (any AgarConfig).  StCfg<...> fixes the sizes of one world at compile time: record offsets and
// the shared-memory map become immediates of the load/store instructions, loops get constant trip counts and
// the branches of absent features (viruses, events, unused channels) disappear.  The host launches a StCfg
// instantiation when the handle's config matches one (the BASELINE worlds), else the RtCfg one; the source
// is the same, so results are identical (tests/test_gpu_parity.py runs both).
struct RtCfg {
    static constexpr bool is_static = false;
    static constexpr int A = 0, BOTS = 0, P = 0, V = 0, F = 0, G = 0, FLAGS = 0, TICKS = 0, K = 0, KC = 0, C = 0;
    struct Layout {};
};
__host__ __device__ constexpr int cpad4(int n) { return (n + 3) & ~3; }
template <int A_, int BOTS_, int P_, int V_, int F_, int G_, int FLAGS_, int TICKS_>
struct StCfg {
    static constexpr bool is_static = true;
    static constexpr int A = A_, BOTS = BOTS_, P = P_, V = V_, F = F_, G = G_, FLAGS = FLAGS_, TICKS = TICKS_;
    static constexpr int K = A_ + BOTS_, KC = K * AGAR_MAX_CELLS;
    static constexpr int C = ((FLAGS_ >> 0) & 1) + ((FLAGS_ >> 1) & 1) + ((FLAGS_ >> 2) & 1) + ((FLAGS_ >> 3) & 1) + ((FLAGS_ >> 4) & 1);
    struct Layout {  // agar_spec_layout (include/agar_spec.h) as constants; the host checks them against the handle's layout
        static constexpr int P_pad = cpad4(P_), V_pad = cpad4(V_), F_pad = cpad4(F_), A_pad = cpad4(A_);
        static constexpr int num_players = K, cells_total = KC;
        static constexpr int pellet_x = 0, pellet_y = P_pad, virus_x = 2 * P_pad, virus_y = 2 * P_pad + V_pad;
        static constexpr int dyn_begin = 2 * P_pad + 2 * V_pad;
        static constexpr int cell_x = dyn_begin, cell_y = cell_x + KC, cell_ix = cell_y + KC, cell_iy = cell_ix + KC;
        static constexpr int cell_m = cell_iy + KC, cell_t = cell_m + KC;
        static constexpr int food_x = cell_t + KC, food_y = food_x + F_pad, food_vx = food_y + F_pad, food_vy = food_vx + F_pad;
        static constexpr int agent_done = food_vy + F_pad, tick = agent_done + A_pad, episode_steps = tick + 1;
        static constexpr int dyn_end = tick + 4, words = dyn_end;
    };
};
template <class CV>
__device__ __forceinline__ decltype(auto) layout_of(const DevParams& p) {
    if constexpr (CV::is_static) return typename CV::Layout{};
    else return (p.L);
}
#define CV_INT(name, runtime_value) (CV::is_static ? CV::name : (runtime_value))

// ------------------------------------------------------------------------- Philox4x32-10
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t& o0, uint32_t& o1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    o0 = c0; o1 = c1;
}

__device__ __forceinline__ float u01(uint32_t u) { return (float)(u >> 8) * (1.0f / 16777216.0f); }

// Out of line on purpose: random positions are drawn on rare paths (respawns, resets), and one copy of
// the ten Philox rounds keeps the hot path of the step kernel small (the SM's instruction cache is 32 KB).
__device__ __noinline__ float2 rand_pos2(uint32_t slot, uint32_t stream, uint32_t tick, uint32_t gid_lo, uint32_t k0,
                                         uint32_t k1, float arena_size) {
    uint32_t o0, o1;
    philox4x32_10(slot, stream, tick, gid_lo, k0, k1, o0, o1);
    return make_float2(u01(o0) * arena_size, u01(o1) * arena_size);
}
__device__ __forceinline__ void rand_pos(const DevParams& p, uint32_t gid_lo, uint32_t gid_hi, uint32_t slot,
                                         uint32_t stream, uint32_t tick, float& x, float& y) {
    const float2 r = rand_pos2(slot, stream, tick, gid_lo, p.seed_lo, p.seed_hi ^ gid_hi, p.c.arena_size);
    x = r.x; y = r.y;
}

__device__ __forceinline__ float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }

// 16-bit atomic min (compare-and-swap loop; claims are rare so it almost never iterates).
// Returns true for the thread that replaced NONE16, i.e. the first claimer of the slot.
__device__ __forceinline__ bool claim_min16(uint16_t* addr, int g) {
    unsigned short old = *reinterpret_cast<volatile unsigned short*>(addr);
    while ((unsigned)g < (unsigned)old) {
        const unsigned short prev = atomicCAS(reinterpret_cast<unsigned short*>(addr), old, (unsigned short)g);
        if (prev == old) return old == NONE16;
        old = prev;
    }
    return false;
}

// two 16-bit event counters per word (a cell cannot eat more than P <= 8192 pellets in a tick)
__device__ __forceinline__ void cnt_inc(uint32_t* cnt32, int g) { atomicAdd(&cnt32[g >> 1], (g & 1) ? 0x10000u : 1u); }
__device__ __forceinline__ int cnt_get(const uint32_t* cnt32, int g) { return reinterpret_cast<const uint16_t*>(cnt32)[g]; }
__device__ __forceinline__ void cnt_clear(uint32_t* cnt32, int g) { reinterpret_cast<uint16_t*>(cnt32)[g] = 0; }

// ------------------------------------------------------------- TMA bulk copies (1-D) + mbarrier
// The arena record and its persisted pellet hash are contiguous blocks of global memory: one thread
// issues `cp.async.bulk` (UBLKCP) into shared memory, completion is counted in bytes on an mbarrier,
// and the whole CTA waits on its phase.  The rewritten part of the record goes back the same way.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ------------------------------------------------------------- half-warp (one player) helpers
// A player's 16 cell slots map to the 16 lanes of a half-warp: hbase is 0 or 16, slot = lane & 15.
// All 32 lanes of the warp call these together (full-mask shuffles are single SHFL instructions;
// partial masks that differ across the warp compile to a MATCH.ANY sequence per operation), so the
// callers keep their control flow warp-uniform and pass mass 0 for lanes without a cell.
#define FULL 0xffffffffu

// numpy's pairwise add-reduce over the first n (<=16) lanes of the half, see np_sum16 in the oracle.
// Both of numpy's code paths are evaluated without branches and the one for n is selected.
__device__ __forceinline__ float hw_np_sum(int hbase, int slot, float a, int n) {
    float seq = 0.0f;  // n < 8: sequential from 0
#pragma unroll
    for (int i = 0; i < 7; ++i) {
        float v = __shfl_sync(FULL, a, hbase + i);
        if (i < n) seq = seq + v;
    }
    // n >= 8: eight accumulators, tree, remainder
    float hi = __shfl_sync(FULL, a, hbase + ((slot + 8) & 15));
    float r = (n == 16) ? a + hi : a;
    r = r + __shfl_xor_sync(FULL, r, 1);
    r = r + __shfl_xor_sync(FULL, r, 2);
    r = r + __shfl_xor_sync(FULL, r, 4);
    float pw = __shfl_sync(FULL, r, hbase);
#pragma unroll
    for (int i = 8; i < 15; ++i) {
        float v = __shfl_sync(FULL, a, hbase + i);
        if (i < n && n < 16) pw = pw + v;
    }
    return n < 8 ? seq : pw;
}

// features/extractors.py:207-212 `position` in numpy float32 evaluation order; false if no cell
struct Centroid { float x, y; int ok; };
__device__ __noinline__ Centroid hw_centroid_ool(int hbase, int slot, float x, float y, float m) {
    Centroid r; r.x = 0.0f; r.y = 0.0f; r.ok = 0;
    const unsigned alive = (__ballot_sync(FULL, m > 0.0f) >> hbase) & 0xFFFFu;
    const int n = __popc(alive);
    const unsigned src = __fns(alive, 0, slot + 1);
    const int srcl = (slot < n) ? (int)src : 0;
    const float w = __shfl_sync(FULL, m, hbase + srcl);
    const float xs = __shfl_sync(FULL, x, hbase + srcl);
    const float ys = __shfl_sync(FULL, y, hbase + srcl);
    const float den = hw_np_sum(hbase, slot, w, n);
    const float nx = hw_np_sum(hbase, slot, xs * w, n);
    const float ny = hw_np_sum(hbase, slot, ys * w, n);
    if (n == 0) return r;
    r.x = nx / den;
    r.y = ny / den;
    r.ok = 1;
    return r;
}
__device__ __forceinline__ bool hw_centroid(int hbase, int slot, float x, float y, float m, float& cx, float& cy) {
    const Centroid r = hw_centroid_ool(hbase, slot, x, y, m);  // all 32 lanes call together
    if (r.ok) { cx = r.x; cy = r.y; }
    return r.ok != 0;
}

// butterfly 8,4,2,1 sum of the 16 slots (treesum16 in the oracle); result valid in every lane
__device__ __forceinline__ float hw_treesum(float v) {
    v = v + __shfl_xor_sync(FULL, v, 8);
    v = v + __shfl_xor_sync(FULL, v, 4);
    v = v + __shfl_xor_sync(FULL, v, 2);
    v = v + __shfl_xor_sync(FULL, v, 1);
    return v;
}

// ------------------------------------------------------------------------- shared memory map
// Pellet hash: HASH_W x HASH_W uniform grid over the arena, built by counting sort (`sorted` holds the pellet ids
// bin after bin, `hstart` the first slot of each bin's run).  Every run is built with HASH_SLACK free slots
// (0xFFFF) behind its pellets, so the hash is kept current instead of rebuilt: a respawning pellet's old entry is
// turned into a free slot and the pellet takes a free slot of its new bin; only a bin without a free slot forces a
// rebuild (next tick).  Readers skip free slots.  The hash is persisted across steps in a side buffer.
#define HASH_W 16
#define HASH_BINS (HASH_W * HASH_W)
#define HASH_SLACK 2    // free slots per bin at build time
#define HASH_FREE 0xFFFFu
#define HASH_HDR 4      // words of the persisted hash header: {epoch, 0, 0, 0}
#define HITS_CAP 128    // claimed pellets listed per tick

struct Smem {
    float* rec;       // [words] the arena record
    int32_t* reci;    // same, as int32
    // pellet hash, kept through the raster
    int32_t* hstart;  // [HASH_BINS+1] first sorted position of each bin
    uint16_t* sorted; // [sort_slots]  pellet ids in bin order, HASH_FREE = free slot
    int32_t* hhdr;    // [HASH_HDR]    header of the persisted hash record: {epoch, 0, 0, 0}
    uint64_t* mbar;   // mbarrier for the bulk copies (2 words)
    // per-tick scratch
    int32_t* hcur;    // [HASH_BINS]   counting-sort cursors (only while the hash is built)
    uint16_t* pwin;   // [P_pad]       winner cell of each pellet this tick (16-bit CAS-min), NONE16 = unclaimed
                      //               (bin cache while the hash is built)
    uint16_t* hits;   // [HITS_CAP]    pellets claimed this tick; past the capacity the resolver scans pwin
    uint16_t* list;   // [KC] gids of the alive cells (unordered)
    uint32_t* cnt32;  // [KC/2] pellets / foods eaten this tick, two 16-bit counters per word
    uint16_t* eater;  // [KC] lowest-gid cell that eats this one / first virus won (NONE16 between uses)
    float* gain;      // [max(KC, F_pad+4)] (doubles as the free-food-slot list of the feed phase)
    int32_t* vwin;    // [V_pad]
    float *Tx, *Ty, *Cx, *Cy;                        // [K_pad] each
    int32_t *palive, *pmask;                         // [K_pad]
    float *dirx, *diry, *mass0;                      // [A_pad] (agents only)
    int32_t* pact;                                   // [A_pad]
    int32_t* rb;                                     // [96] raster: other players' cells in view (aliases hcur)
    int32_t* sc;                                     // [32] scalars + [32] list-membership bitmask
};
enum { SC_NLIST = 0, SC_ANYFEED, SC_ANYSPLIT, SC_FOODANY, SC_ANYEAT, SC_ANYVHIT, SC_EVCOUNT, SC_RESET, SC_ALLDONE,
       SC_NHITS, SC_HDIRTY, SC_REBUILD, SC_MULTI, SC_HASHOK, SC_NOB, SC_HBUILT, SC_MERGEABLE, SC_COUNT };
static_assert(SC_COUNT <= 32, "scalar slots overlap the list-membership bitmask");

__host__ __device__ inline size_t pad4z(size_t n) { return (n + 3) & ~(size_t)3; }
template <class LT>
__host__ __device__ inline size_t gain_words(const LT& L) {
    const size_t a = (size_t)L.cells_total, b = (size_t)L.F_pad + 4;
    return a > b ? a : b;
}
template <class LT>
__host__ __device__ inline size_t scratch_words(const LT& L) {
    // hcur | pwin | hits | list, eater (16-bit) | cnt32 | gain
    return HASH_BINS + (size_t)L.P_pad / 2 + HITS_CAP / 2 + (size_t)L.cells_total + (size_t)L.cells_total / 2 + gain_words(L);
}
template <class LT>
__host__ __device__ inline size_t union_words(const LT& L, int G) {
    (void)G;  // the observation needs no shared-memory tile: it is written from registers and patched in place
    return pad4z(scratch_words(L));
}
template <class LT>
__host__ __device__ inline size_t kpad(const LT& L) { return pad4z((size_t)L.num_players); }

// slots of `sorted`: every pellet plus the slack of every bin (a multiple of 8 so that the words divide by 4)
template <class LT>
__host__ __device__ inline size_t sort_slots(const LT& L) {
    return ((size_t)L.P_pad + HASH_SLACK * HASH_BINS + 7) & ~(size_t)7;
}
// persisted hash record: the shared-memory block [hstart | sorted] followed by the header
template <class LT>
__host__ __device__ inline size_t hash_body_words(const LT& L) {
    return HASH_BINS + 4 + sort_slots(L) / 2;
}
template <class LT>
__host__ __device__ inline size_t hash_record_words(const LT& L) { return hash_body_words(L) + HASH_HDR; }

template <class LT>
__host__ __device__ inline size_t smem_bytes(const LT& L, int G) {
    size_t w = 0;
    w += (size_t)L.words;                                   // rec
    w += hash_body_words(L) + HASH_HDR + 4;                 // hash kept, header, mbarrier
    w += union_words(L, G);                                 // per-tick scratch
    w += (size_t)(L.V_pad + 4);                             // vwin
    w += kpad(L) * 6 + (size_t)L.A_pad * 4;                 // per-player and per-agent arrays
    w += 64;                                                // scalars + alive-list membership bits
    return w * 4 + 16;
}

template <class LT>
__device__ __forceinline__ Smem carve(float* base, const LT& L, int G) {
    Smem s;
    float* p = base;
    s.rec = p; s.reci = (int32_t*)p; p += L.words;           // words % 4 == 0
    s.hstart = (int32_t*)p; p += HASH_BINS + 4;
    s.sorted = (uint16_t*)p; p += sort_slots(L) / 2;
    s.hhdr = (int32_t*)p; p += HASH_HDR;                     // contiguous with the body: one bulk copy
    s.mbar = (uint64_t*)p; p += 4;
    {
        int32_t* h = (int32_t*)p;
        s.hcur = h; s.rb = h; h += HASH_BINS;                // rb: raster scratch, used when no hash is being built
        s.pwin = (uint16_t*)h; h += L.P_pad / 2;
        s.hits = (uint16_t*)h; h += HITS_CAP / 2;
        s.list = (uint16_t*)h; h += L.cells_total / 2;
        s.eater = (uint16_t*)h; h += L.cells_total / 2;
        s.cnt32 = (uint32_t*)h; h += L.cells_total / 2;
        s.gain = (float*)h;
    }
    p += union_words(L, G);
    s.vwin = (int32_t*)p; p += L.V_pad + 4;                  // V_pad % 4 == 0
    const size_t kp = kpad(L);
    s.Tx = p; p += kp; s.Ty = p; p += kp; s.Cx = p; p += kp; s.Cy = p; p += kp;
    s.palive = (int32_t*)p; p += kp; s.pmask = (int32_t*)p; p += kp;
    s.dirx = p; p += L.A_pad; s.diry = p; p += L.A_pad; s.mass0 = p; p += L.A_pad;
    s.pact = (int32_t*)p; p += L.A_pad;
    s.sc = (int32_t*)p; p += 64;
    return s;
}

// -------------------------------------------------------------------------------- events
__device__ __noinline__ void ev_push_slow(int* counter, int32_t* events, int cap, int arena, int tick, int type, int x, int y) {
    int i = atomicAdd(counter, 1);
    if (i < cap) *reinterpret_cast<int4*>(events + ((size_t)arena * cap + i) * 4) = make_int4(tick, type, x, y);
}
template <class CV>
__device__ __forceinline__ void ev_push(const DevParams& p, const KArgs& a, const Smem& s, int arena, int tick,
                                        int type, int x, int y) {
    if (!CV::is_static && p.c.event_capacity > 0) ev_push_slow(&s.sc[SC_EVCOUNT], a.events, p.c.event_capacity, arena, tick, type, x, y);
}

// ----------------------------------------------------------------------------- sub-phases
template <class CV>
__device__ __forceinline__ void clear_cell(const DevParams& p, const Smem& s, int g) {
    decltype(auto) L = layout_of<CV>(p);
    s.rec[L.cell_x + g] = 0.0f; s.rec[L.cell_y + g] = 0.0f; s.rec[L.cell_ix + g] = 0.0f;
    s.rec[L.cell_iy + g] = 0.0f; s.rec[L.cell_m + g] = 0.0f; s.rec[L.cell_t + g] = 0.0f;
}

// env.reset() for one arena, in shared memory (reset_arena in the oracle)
// (out of line, re-deriving the shared-memory map itself: rare path, kept off the hot instruction stream)
template <class CV>
__device__ __noinline__ void reset_arena(const DevParams& p, uint32_t gid_lo, uint32_t gid_hi) {
    extern __shared__ __align__(16) float smem_raw[];
    decltype(auto) L = layout_of<CV>(p);
    const Smem s = carve(smem_raw, L, CV_INT(G, p.c.grid_size));
    const int P = CV_INT(P, p.c.num_pellets), V = CV_INT(V, p.c.num_viruses), K = CV_INT(K, p.K);
    const int tid = threadIdx.x, nt = blockDim.x;
    uint32_t tick = (uint32_t)s.reci[L.tick];
    __syncthreads();
    for (int i = tid; i < L.words; i += nt) s.rec[i] = 0.0f;
    __syncthreads();
    if (tid == 0) s.reci[L.tick] = (int32_t)tick;
    for (int i = tid; i < L.P_pad; i += nt) {
        float x = -1.0f, y = -1.0f;
        if (i < P) rand_pos(p, gid_lo, gid_hi, (uint32_t)i, AGAR_STREAM_PELLET_INIT, tick, x, y);
        s.rec[L.pellet_x + i] = x; s.rec[L.pellet_y + i] = y;
    }
    for (int i = tid; i < L.V_pad; i += nt) {
        float x = -1.0f, y = -1.0f;
        if (i < V) rand_pos(p, gid_lo, gid_hi, (uint32_t)i, AGAR_STREAM_VIRUS_INIT, tick, x, y);
        s.rec[L.virus_x + i] = x; s.rec[L.virus_y + i] = y;
    }
    for (int i = tid; i < L.F_pad; i += nt) s.rec[L.food_x + i] = -1.0f;
    for (int k = tid; k < K; k += nt) {
        float x, y;
        rand_pos(p, gid_lo, gid_hi, (uint32_t)k, AGAR_STREAM_PLAYER_SPAWN, tick, x, y);
        s.rec[L.cell_x + k * MC] = x; s.rec[L.cell_y + k * MC] = y;
        s.rec[L.cell_m + k * MC] = p.c.start_mass;
    }
    __syncthreads();
}

// pellet bin: monotone in the coordinate, so a conservative coordinate range maps to a bin range
__device__ __forceinline__ int hash_coord(const DevParams& p, float v) {
    int b = (int)(v * p.hash_inv);
    return b < 0 ? 0 : (b > HASH_W - 1 ? HASH_W - 1 : b);
}

// counting sort of the present pellets into the HASH_W x HASH_W grid, HASH_SLACK free slots behind every bin's
// pellets (all threads; ends synchronised).
// Order inside a bin is arbitrary: every consumer reduces with min / integer counts.
// Out of line: with the hash persisted across steps this runs only after resets and respawn overflows.
template <class CV>
__device__ __noinline__ void build_hash(const DevParams& p) {
    extern __shared__ __align__(16) float smem_raw[];
    decltype(auto) L = layout_of<CV>(p);
    const Smem s = carve(smem_raw, L, CV_INT(G, p.c.grid_size));
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, P = CV_INT(P, p.c.num_pellets);
    const float* px = s.rec + L.pellet_x; const float* py = s.rec + L.pellet_y;
    for (int i = tid; i < HASH_BINS; i += nt) s.hcur[i] = 0;
    for (int i = tid; i < (int)sort_slots(L) / 2; i += nt) reinterpret_cast<uint32_t*>(s.sorted)[i] = 0xFFFFFFFFu;  // all free
    __syncthreads();
    for (int i = tid; i < P; i += nt) {
        const float x = px[i];
        int b = 0xFFFF;
        if (x >= 0.0f) { b = hash_coord(p, py[i]) * HASH_W + hash_coord(p, x); atomicAdd(&s.hcur[b], 1); }
        s.pwin[i] = (uint16_t)b;  // bin cache for the scatter pass (which restores NONE16 == 0xFFFF)
    }
    __syncthreads();
    if (tid < 32) {
        constexpr int PER = HASH_BINS / 32;
        int c[PER], tot = 0;
#pragma unroll
        for (int q = 0; q < PER; ++q) { c[q] = s.hcur[lane * PER + q] + HASH_SLACK; tot += c[q]; }  // run = pellets + slack
        int incl = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        int run = incl - tot;
#pragma unroll
        for (int q = 0; q < PER; ++q) { s.hstart[lane * PER + q] = run; s.hcur[lane * PER + q] = run; run += c[q]; }
        if (lane == 31) s.hstart[HASH_BINS] = incl;
        if (lane == 0) { s.sc[SC_REBUILD] = 0; s.sc[SC_NHITS] = 0; s.sc[SC_HASHOK] = 1; s.sc[SC_HBUILT] = 1; }
    }
    __syncthreads();
    for (int i = tid; i < P; i += nt) {
        const int b = s.pwin[i];
        if (b != 0xFFFF) { s.sorted[atomicAdd(&s.hcur[b], 1)] = (uint16_t)i; s.pwin[i] = NONE16; }
    }
    __syncthreads();
}

// A pellet left position (x0, y0): its entry in that bin's run becomes a free slot (one thread per pellet).
__device__ __forceinline__ void hash_remove(const DevParams& p, const Smem& s, int i, float x0, float y0) {
    const int b = hash_coord(p, y0) * HASH_W + hash_coord(p, x0);
    for (int j = s.hstart[b], j1 = s.hstart[b + 1]; j < j1; ++j)
        if (s.sorted[j] == (uint16_t)i) { s.sorted[j] = (uint16_t)HASH_FREE; return; }
}
// ... and appeared at (x1, y1): it takes a free slot of that bin's run; false if the bin has none (rebuild needed).
// Several threads may insert into one bin at once: slots are claimed with a 16-bit compare-and-swap.
__device__ __forceinline__ bool hash_insert(const DevParams& p, const Smem& s, int i, float x1, float y1) {
    const int b = hash_coord(p, y1) * HASH_W + hash_coord(p, x1);
    for (int j = s.hstart[b], j1 = s.hstart[b + 1]; j < j1; ++j)
        if (s.sorted[j] == (uint16_t)HASH_FREE &&
            atomicCAS(reinterpret_cast<unsigned short*>(&s.sorted[j]), (unsigned short)HASH_FREE, (unsigned short)i) == (unsigned short)HASH_FREE)
            return true;
    return false;
}

// all threads: unordered list of the alive cells' gids, per-player alive masks, "some player has several cells" and
// "a merge is possible" flags, list-membership bits.  SC_NLIST, SC_MULTI, SC_MERGEABLE must be 0 on entry (a barrier
// earlier).  One thread takes four cell slots (128-bit loads of mass and merge timer): the four threads of a
// player assemble its 16-bit masks by shuffles, eight threads a 32-bit membership word.
template <class CV>
__device__ __forceinline__ void build_list(const DevParams& p, const Smem& s) {
    decltype(auto) L = layout_of<CV>(p);
    const int KC = CV_INT(KC, p.KC);
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
    const float4* m4 = reinterpret_cast<const float4*>(s.rec + L.cell_m);   // segments are 16-byte aligned, KC % 16 == 0
    const float4* t4 = reinterpret_cast<const float4*>(s.rec + L.cell_t);
    for (int base = 0; base < KC / 4; base += nt) {
        const int q = base + tid;          // slots 4q .. 4q+3
        const bool valid = q < KC / 4;
        const float4 m = valid ? m4[q] : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 t = valid ? t4[q] : make_float4(1.f, 1.f, 1.f, 1.f);
        const unsigned nib = (m.x > 0.0f ? 1u : 0u) | (m.y > 0.0f ? 2u : 0u) | (m.z > 0.0f ? 4u : 0u) | (m.w > 0.0f ? 8u : 0u);
        const unsigned rdy = ((m.x > 0.0f && t.x == 0.0f) ? 1u : 0u) | ((m.y > 0.0f && t.y == 0.0f) ? 2u : 0u) |
                             ((m.z > 0.0f && t.z == 0.0f) ? 4u : 0u) | ((m.w > 0.0f && t.w == 0.0f) ? 8u : 0u);
        // membership word of 32 slots = the nibbles of 8 consecutive threads; a player's masks = 4 consecutive threads
        unsigned w = (nib | (rdy << 16)) << (4 * (lane & 3));   // low half: alive nibble at its place, high half: ready nibble
        w |= __shfl_xor_sync(FULL, w, 1);
        w |= __shfl_xor_sync(FULL, w, 2);                        // 16 alive bits | 16 ready bits of the player, in its 4 lanes
        unsigned mw = (w & 0xFFFFu) << (16 * ((lane >> 2) & 1));
        mw |= __shfl_xor_sync(FULL, mw, 4);                      // membership word in the 8 lanes
        if (valid && (lane & 7) == 0) s.sc[32 + (q >> 3)] = (int)mw;
        if (valid && (lane & 3) == 0) {
            s.pmask[q >> 2] = (int)(w & 0xFFFFu);
            if (__popc(w & 0xFFFFu) >= 2) s.sc[SC_MULTI] = 1;
            if (__popc(w >> 16) >= 2) s.sc[SC_MERGEABLE] = 1;    // the merge phase has work
        }
        // list: warp prefix of the per-thread counts, one atomic per warp
        const int cnt = __popc(nib);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += v;
        }
        const int tot = __shfl_sync(FULL, incl, 31);
        if (tot) {
            int pos = 0;
            if (lane == 31) pos = atomicAdd(&s.sc[SC_NLIST], tot);
            pos = __shfl_sync(FULL, pos, 31) + incl - cnt;
            if (nib & 1u) s.list[pos++] = (uint16_t)(4 * q);
            if (nib & 2u) s.list[pos++] = (uint16_t)(4 * q + 1);
            if (nib & 4u) s.list[pos++] = (uint16_t)(4 * q + 2);
            if (nib & 8u) s.list[pos++] = (uint16_t)(4 * q + 3);
        }
    }
}

// features/extractors.py:77-78: grid index of an offset, int() truncation toward zero
// (box is a power of two in every BASELINE config: the product is then exact and the IEEE division,
// some 20 instructions plus a slow path at every site, is kept out of line)
__device__ __noinline__ float div_rn(float a, float b) { return a / b; }
__device__ __forceinline__ int grid_bin(const DevParams& p, float d, int half) {
    return (int)(p.box_inv != 0.0f ? d * p.box_inv : div_rn(d, p.box)) + half;
}

// One warp: given up to 32 in-view cell entities (key = entity order, bin, mass; bin < 0 = none)
// one per lane, leaves in `val` the sum of the masses of all entities sharing the lane's bin,
// accumulated in ascending key order (bit-identical to the sequential `+=` of add_ones,
// features/extractors.py:75-82) and returns true for the lane that holds the last entity of its bin.
__device__ __noinline__ float2 ordered_bin_sums_ool(int key, int bin, float m, int cnt, int lane) {
    int rank = 0;
    for (int i = 0; i < cnt; ++i) {
        int ki = __shfl_sync(0xffffffffu, key, i);
        rank += (ki < key) ? 1 : 0;
    }
    float acc = 0.0f;
    bool later = false;
    for (int r = 0; r < cnt; ++r) {
        unsigned bal = __ballot_sync(0xffffffffu, lane < cnt && rank == r);
        int src = __ffs(bal) - 1;
        int br = __shfl_sync(0xffffffffu, bin, src);
        float mr = __shfl_sync(0xffffffffu, m, src);
        if (br == bin && r <= rank) acc = acc + mr;
        if (br == bin && r > rank) later = true;
    }
    return make_float2(acc, (lane < cnt && bin >= 0 && !later) ? 1.0f : 0.0f);
}
__device__ __forceinline__ bool ordered_bin_sums(int key, int bin, float m, int cnt, int lane, float& val) {
    const float2 r = ordered_bin_sums_ool(key, bin, m, cnt, lane);  // all 32 lanes call together
    val = r.x;
    return r.y != 0.0f;
}

// features/extractors.py:84-96: the grid cell with index i on an axis is out of bounds when its world
// corner (i - G//2)*box + c lies outside [0, arena)
__device__ __forceinline__ bool oob_cell(const DevParams& p, int i, float c) {
    const float w = p.oob_off[i] + c;
    return !(0.0f <= w && w < p.c.arena_size);
}

// One count entity (value 1, add_ones with len(entity) <= 2) of an agent whose centroid is (cx, cy):
// float adds of 1.0 are exact, so a global float reduction gives the reference's count in any order.
// Cells that the sentinel pass overwrites afterwards (features/extractors.py:71-73) keep their -1.
__device__ __forceinline__ void count_entity(const DevParams& p, float* __restrict__ chan, float x, float y, float cx,
                                             float cy, int G, int half, float lim) {
    const float dx = x - cx, dy = y - cy;
    if (fabsf(dx) > lim || fabsf(dy) > lim) return;  // cannot land in the grid
    const int gx = grid_bin(p, dx, half), gy = grid_bin(p, dy, half);
    if (0 <= gx && gx < G && 0 <= gy && gy < G && !oob_cell(p, gx, cx) && !oob_cell(p, gy, cy))
        atomicAdd(chan + gx * G + gy, 1.0f);
}

// GridFeatureExtractor.extract for every agent of the arena held in s.rec
// (features/extractors.py:39-96; oracle_grid_extract).  obs_arena points at [A,C,G,G].
// use_hash: the pellet hash of this step is valid (pellets are then taken from the bins under
// the view window instead of scanning all of them).  nlist >= 0: the alive list of this step is valid.
//
// Every channel of an agent is the same base image — zeros, with -1 in the out-of-bounds rows and
// columns — plus a few occupied bins.  The base image goes out from registers with 128-bit
// stores (one pass writes all channels; plain write-back stores, not streaming ones: the lines must
// still be in L2 when their patches arrive or DRAM sees read-modify-writes); after a barrier the
// occupied bins are patched: count
// channels (pellets, viruses, foods) by float reductions of 1.0 in global memory, mass channels (own
// cells, others) by warp 0 with sums taken in entity order.
template <class CV>
__device__ void raster_arena(const DevParams& p, const Smem& s, float* __restrict__ obs_arena, float* __restrict__ meta_arena,
                             bool use_hash, int nlist) {
    decltype(auto) L = layout_of<CV>(p);
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
    const int hbase = lane & 16, slot = lane & 15;
    const int G = CV_INT(G, p.c.grid_size), GG = G * G, A = CV_INT(A, p.c.num_agents), K = CV_INT(K, p.K), C = CV_INT(C, p.C);
    const int KC = CV_INT(KC, p.KC), obs_flags = CV_INT(FLAGS, p.c.obs_flags);
    const float lim = p.raster_lim;
    const int half = G / 2;
    const float* px = s.rec + L.pellet_x; const float* py = s.rec + L.pellet_y;
    const float* cm = s.rec + L.cell_m; const float* cxs = s.rec + L.cell_x; const float* cys = s.rec + L.cell_y;
    // other players' cells in view (unordered): key/bin/mass [32] each
    int* ob_key = s.rb; int* ob_bin = s.rb + 32; float* ob_m = (float*)(s.rb + 64);

    // channel numbers and which channels carry the sentinel
    int ch_pel = -1, ch_cells = -1, ch_others = -1, ch_vir = -1, ch_food = -1, sentmask = 0;
    {
        int ch = 0;
        if (obs_flags & AGAR_OBS_PELLETS) { ch_pel = ch; sentmask |= 1 << ch; ++ch; }
        if (obs_flags & AGAR_OBS_CELLS) { ch_cells = ch; sentmask |= 1 << ch; ++ch; }
        if (obs_flags & AGAR_OBS_OTHERS) { ch_others = ch; if (K > 1) sentmask |= 1 << ch; ++ch; }  // no pass without others
        if (obs_flags & AGAR_OBS_VIRUSES) { ch_vir = ch; sentmask |= 1 << ch; ++ch; }
        if (obs_flags & AGAR_OBS_FOODS) { ch_food = ch; sentmask |= 1 << ch; ++ch; }
    }
    // centroids of all agents (half-warp per agent; single-cell agents take the one-term form)
    for (int base = (tid & ~31); base < A * MC; base += nt) {  // warp-uniform; warps without a valid lane skip
        const int g = base + lane;
        const bool valid = g < A * MC;
        const float m = valid ? cm[g] : 0.0f, x = valid ? cxs[g] : 0.0f, y = valid ? cys[g] : 0.0f;
        const int n = __popc((__ballot_sync(FULL, m > 0.0f) >> hbase) & 0xFFFFu);
        if (!__any_sync(FULL, n >= 2)) {
            if (valid && n == 0 && slot == 0) s.palive[g >> 4] = 0;
            if (m > 0.0f) { s.Cx[g >> 4] = (x * m) / m; s.Cy[g >> 4] = (y * m) / m; s.palive[g >> 4] = 1; }
        } else {
            float cx = 0.0f, cy = 0.0f;
            const bool has = hw_centroid(hbase, slot, x, y, m, cx, cy);
            if (valid && slot == 0) { s.Cx[g >> 4] = cx; s.Cy[g >> 4] = cy; s.palive[g >> 4] = has ? 1 : 0; }
        }
    }
    if (tid == 0) s.sc[SC_NOB] = 0;
    PHASE(17);  // raster: setup + centroids
    // write-out geometry: when 4*nt is a multiple of G a thread keeps its 4 columns and walks rows
    const bool fast = ((4 * nt) % G) == 0;
    const int gy_t = (4 * tid) % G, gx_t = (4 * tid) / G, rstep = (4 * nt) / G;
    const float4 zero4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f), neg4 = make_float4(-1.0f, -1.0f, -1.0f, -1.0f);
    __syncthreads();
    if (meta_arena)  // what the host needs to rebuild an agent's sentinel rows / columns (agar_step_host)
        for (int a = tid; a < A; a += nt)
            *reinterpret_cast<float4*>(meta_arena + 4 * a) = make_float4(s.Cx[a], s.Cy[a], s.palive[a] ? 1.0f : 0.0f, 0.0f);

    for (int a = 0; a < A; ++a) {
        const float cx = s.Cx[a], cy = s.Cy[a];
        const bool has = s.palive[a] != 0;
        float* out_a = obs_arena + (size_t)a * C * GG;
        // ---- gather: other players' cells in view ----
        if (has && ch_others >= 0) {
            // after a step the alive list names every cell that can exist (dead entries have mass 0); without one
            // (observe-only calls, an arena just reset) every slot is examined
            for (int q = tid; q < (nlist >= 0 ? nlist : KC); q += nt) {
                const int g = nlist >= 0 ? (int)s.list[q] : q;
                const float m = cm[g];
                if (!(m > 0.0f) || (g >> 4) == a) continue;
                const float dx = cxs[g] - cx, dy = cys[g] - cy;
                if (fabsf(dx) > lim || fabsf(dy) > lim) continue;
                const int gx = grid_bin(p, dx, half), gy = grid_bin(p, dy, half);
                if (0 <= gx && gx < G && 0 <= gy && gy < G) {
                    const int e = atomicAdd(&s.sc[SC_NOB], 1);
                    if (e < 32) { ob_key[e] = g; ob_bin[e] = gx * G + gy; ob_m[e] = m; }
                }
            }
        }
        PHASE(18);  // raster: barrier + gather others
        // ---- base image of every channel ----
        {
            float4* out4 = reinterpret_cast<float4*>(out_a);
            const int cstride = GG / 4;
            const int smask = has ? sentmask : 0;  // an agent without cells observes zeros
            if (fast) {
                float4 row = zero4;
                if (smask) {
                    if (oob_cell(p, gy_t, cy)) row.x = -1.0f;
                    if (oob_cell(p, gy_t + 1, cy)) row.y = -1.0f;
                    if (oob_cell(p, gy_t + 2, cy)) row.z = -1.0f;
                    if (oob_cell(p, gy_t + 3, cy)) row.w = -1.0f;
                }
#pragma unroll 1  // measured: smaller code beats the unrolled store loop (cfg2 +1 %, cfg1 +1.8 %)
                for (int i = tid, gx = gx_t; i < cstride; i += nt, gx += rstep) {
                    const float4 v = (smask && oob_cell(p, gx, cx)) ? neg4 : row;
                    for (int ch = 0; ch < C; ++ch) out4[ch * cstride + i] = ((smask >> ch) & 1) ? v : zero4;
                }
            } else {
                for (int i = tid; i < cstride; i += nt) {
                    float4 v = zero4;
                    if (smask) {
                        const int cell = i * 4, gx = cell / G, gy = cell - gx * G;
                        if (oob_cell(p, gx, cx)) v = neg4;
                        else {
                            if (oob_cell(p, gy, cy)) v.x = -1.0f;
                            if (oob_cell(p, gy + 1, cy)) v.y = -1.0f;
                            if (oob_cell(p, gy + 2, cy)) v.z = -1.0f;
                            if (oob_cell(p, gy + 3, cy)) v.w = -1.0f;
                        }
                    }
                    for (int ch = 0; ch < C; ++ch) out4[ch * cstride + i] = ((smask >> ch) & 1) ? v : zero4;
                }
            }
        }
        PHASE(19);  // raster: base image stores
        __syncthreads();  // the base image is ordered before the patches; the gather list is complete
        PHASE(20);  // raster: barrier after the stores
        if (has) {
            const int nob = s.sc[SC_NOB];
            // ---- warp 0: own cells and others, sums in entity order, stored over the base image ----
            if (tid < 32) {
                if (ch_cells >= 0) {
                    const int g = a * MC + lane;
                    const float m = (lane < MC) ? cm[g] : 0.0f;
                    int bin = -1;
                    if (m > 0.0f) {
                        const float dx = cxs[g] - cx, dy = cys[g] - cy;
                        if (!(fabsf(dx) > lim || fabsf(dy) > lim)) {
                            const int gx = grid_bin(p, dx, half), gy = grid_bin(p, dy, half);
                            if (0 <= gx && gx < G && 0 <= gy && gy < G) bin = gx * G + gy;
                        }
                    }
                    const unsigned inb = __ballot_sync(FULL, bin >= 0);
                    float val = m;  // a single cell in view: 0 + m
                    bool last = bin >= 0;
                    // entities in distinct bins (the usual case): every sum is its single term
                    const unsigned same = __match_any_sync(FULL, bin);
                    if (__popc(inb) > 1 && __any_sync(FULL, bin >= 0 && same != (1u << lane)))
                        last = ordered_bin_sums(lane, bin, m, MC, lane, val);
                    if (last) {
                        const int gx = bin / G, gy = bin - gx * G;
                        if (!(oob_cell(p, gx, cx) || oob_cell(p, gy, cy))) out_a[(size_t)ch_cells * GG + bin] = val;
                    }
                }
                if (ch_others >= 0 && nob > 0 && nob <= 32) {
                    const int key = lane < nob ? ob_key[lane] : NONE_I;
                    const int bin = lane < nob ? ob_bin[lane] : -1;
                    const float m = lane < nob ? ob_m[lane] : 0.0f;
                    float val = m;
                    bool last = lane < nob;
                    const unsigned same = __match_any_sync(FULL, bin);
                    if (nob > 1 && __any_sync(FULL, lane < nob && same != (1u << lane)))
                        last = ordered_bin_sums(key, bin, m, nob, lane, val);
                    if (last) {
                        const int gx = bin / G, gy = bin - gx * G;
                        if (!(K > 1 && (oob_cell(p, gx, cx) || oob_cell(p, gy, cy)))) out_a[(size_t)ch_others * GG + bin] = val;
                    }
                }
                if (ch_others >= 0 && nob > 32 && lane == 0) {
                    // crowded view (more than 32 other cells inside it): one thread adds the masses in (player,
                    // slot) order onto the zeros stored above — the same float sequence as the reference's `+=`
                    float* outc = out_a + (size_t)ch_others * GG;
                    for (int g = 0; g < KC; ++g) {
                        const float m = cm[g];
                        if (!(m > 0.0f) || (g >> 4) == a) continue;
                        const float dx = cxs[g] - cx, dy = cys[g] - cy;
                        if (fabsf(dx) > lim || fabsf(dy) > lim) continue;
                        const int gx = grid_bin(p, dx, half), gy = grid_bin(p, dy, half);
                        if (0 <= gx && gx < G && 0 <= gy && gy < G && !(K > 1 && (oob_cell(p, gx, cx) || oob_cell(p, gy, cy)))) {
                            volatile float* q = outc + gx * G + gy;
                            *q = *q + m;
                        }
                    }
                }
            }
            PHASE(21);  // raster: mass patches (warp 0)
            // ---- count channels ----
            if (ch_pel >= 0) {
                float* chan = out_a + (size_t)ch_pel * GG;
                if (use_hash) {
                    // pellets under the view window: the hash rows
                    const int bx0 = hash_coord(p, cx - lim), bx1 = hash_coord(p, cx + lim);
                    const int by0 = hash_coord(p, cy - lim), by1 = hash_coord(p, cy + lim);
                    // warp 0 is busy with the mass patches above: the other warps share the hash rows
                    const int nw = (nt >> 5) > 1 ? (nt >> 5) - 1 : 1, wi = (nt >> 5) > 1 ? (tid >> 5) - 1 : 0;
                    for (int by = by0 + wi; wi >= 0 && by <= by1; by += nw) {
                        const int j1 = s.hstart[by * HASH_W + bx1 + 1];
                        for (int j = s.hstart[by * HASH_W + bx0] + lane; j < j1; j += 32) {
                            const int i = s.sorted[j];
                            if (i == (int)HASH_FREE) continue;
                            count_entity(p, chan, px[i], py[i], cx, cy, G, half, lim);  // every present pellet has exactly one entry
                        }
                    }
                } else {
                    for (int i = tid; i < CV_INT(P, p.c.num_pellets); i += nt) {
                        const float x = px[i];
                        if (x >= 0.0f) count_entity(p, chan, x, py[i], cx, cy, G, half, lim);
                    }
                }
            }
            if (ch_vir >= 0) {
                float* chan = out_a + (size_t)ch_vir * GG;
                const float* ex = s.rec + L.virus_x; const float* ey = s.rec + L.virus_y;
                for (int i = tid; i < CV_INT(V, p.c.num_viruses); i += nt) count_entity(p, chan, ex[i], ey[i], cx, cy, G, half, lim);
            }
            if (ch_food >= 0) {
                float* chan = out_a + (size_t)ch_food * GG;
                const float* ex = s.rec + L.food_x; const float* ey = s.rec + L.food_y;
                for (int i = tid; i < CV_INT(F, p.c.max_foods); i += nt) {
                    const float x = ex[i];
                    if (x >= 0.0f) count_entity(p, chan, x, ey[i], cx, cy, G, half, lim);
                }
            }
        }
        if (a + 1 < A) {  // the gather list and its counter are reused by the next agent
            __syncthreads();
            if (tid == 0) s.sc[SC_NOB] = 0;
            __syncthreads();
        }
    }
}

// Virus pops of one tick: a cell pops on the lowest virus it won; pops inside a
// player are sequential in slot order.  All threads; ends synchronised.  Out of line: rare.
template <class CV>
__device__ __noinline__ void virus_pops(const DevParams& p, const KArgs& a, int arena, int tk, uint32_t tick,
                                        uint32_t gid_lo, uint32_t gid_hi) {
    extern __shared__ __align__(16) float smem_raw[];
    decltype(auto) L = layout_of<CV>(p);
    const AgarConfig& c = p.c;
    const Smem s = carve(smem_raw, L, CV_INT(G, c.grid_size));
    const int tid = threadIdx.x, nt = blockDim.x;
    const int K = CV_INT(K, p.K), V = CV_INT(V, c.num_viruses);
    float* cx = s.rec + L.cell_x; float* cy = s.rec + L.cell_y; float* cix = s.rec + L.cell_ix;
    float* ciy = s.rec + L.cell_iy; float* cm = s.rec + L.cell_m; float* ct = s.rec + L.cell_t;
    float* vx = s.rec + L.virus_x; float* vy = s.rec + L.virus_y;
    float* grec = a.state + (size_t)arena * L.words;
    int32_t* inl = s.sc + 32;
    const int n = s.sc[SC_NLIST];
    for (int q = tid; q < n; q += nt) {
        const int g = s.list[q];
        s.eater[g] = NONE16;  // the speculative eat test is void: masses change below
        if (!(cm[g] >= c.virus_pop_mass)) continue;
        for (int v = 0; v < V; ++v)
            if (s.vwin[v] == g) { s.eater[g] = (uint16_t)v; break; }   // eater[] doubles as "first virus won"
    }
    __syncthreads();
    // one thread per player walks its slots in order (free slots are consumed sequentially
    // inside a player, exactly as the oracle does)
    for (int k = tid; k < K; k += nt) {
#pragma unroll 1
        for (int sl = 0; sl < MC; ++sl) {
            const int g = k * MC + sl;
            const int v = s.eater[g];
            if (v == NONE16) continue;
            s.eater[g] = NONE16;
            float m1 = cm[g] + c.virus_mass;
            int nfree = 0;
#pragma unroll 1
            for (int q = 0; q < MC; ++q) nfree += (cm[k * MC + q] == 0.0f) ? 1 : 0;
            int np = nfree < c.pop_pieces ? nfree : c.pop_pieces;
            ev_push<CV>(p, a, s, arena, tk, AGAR_EV_VIRUS_POP, g, v);
            if (np > 0) {
                const float DX[16] = AGAR_DIR16_X; const float DY[16] = AGAR_DIR16_Y;
                float keep = m1 * 0.5f;
                float piece = (m1 - keep) / (float)np;
                cm[g] = keep; ct[g] = c.merge_delay;
                const float gx = cx[g], gy = cy[g];
                int q = 0;
#pragma unroll 1
                for (int j = 0; j < np; ++j) {
                    while (cm[k * MC + q] != 0.0f) ++q;  // j-th free slot in ascending order
                    int d = k * MC + q;
                    cx[d] = gx; cy[d] = gy;
                    cix[d] = DX[j] * c.pop_speed; ciy[d] = DY[j] * c.pop_speed;
                    cm[d] = piece; ct[d] = c.merge_delay;
                    ev_push<CV>(p, a, s, arena, tk, AGAR_EV_POP_PIECE, g, d);
                    if (!(atomicOr(&inl[d >> 5], 1 << (d & 31)) & (1 << (d & 31)))) s.list[atomicAdd(&s.sc[SC_NLIST], 1)] = (uint16_t)d;
                }
                s.sc[SC_MULTI] = 1;
                if (!(c.merge_delay > 0.0f)) s.sc[SC_MERGEABLE] = 1;
            } else {
                cm[g] = m1;
            }
            float nx, ny;
            rand_pos(p, gid_lo, gid_hi, (uint32_t)v, AGAR_STREAM_VIRUS_RESPAWN, tick, nx, ny);
            vx[v] = nx; vy[v] = ny;
            if (!(a.flags & F_READONLY)) { grec[L.virus_x + v] = nx; grec[L.virus_y + v] = ny; }
        }
    }
    if (tid == 0) { s.sc[SC_ANYVHIT] = 0; s.sc[SC_ANYEAT] = 0; }
    __syncthreads();
}

// --------------------------------------------------------------------------- the kernel
// grid = number of arenas (records); block = kThreads; dynamic smem = smem_bytes().
//
// Phase plan of one step (B = __syncthreads):
//   stage record + persisted hash (TMA bulk copies) | B | [pellet hash rebuild] | alive list + player masks | B |
//   players: respawn, centroid | B | agents: action -> target; bots: nearest pellet through the hash | B |
//   [split] [feed] | B |
//   per tick:  motion + pellet query (four lanes per alive cell, hash rows + respawn list) | B |
//              [resolve claimed pellets: count, respawn by Philox, respawn list | B] | [foods] |
//              apply gains + virus claims, cell-eats-cell test | B | [pops, test again] | [ordered gains] |
//              [merge (half-warp per player, shuffles) | B]
//   rewards/dones | B | [auto reset] | persist hash | write back (bulk store) | raster.
// kMinBlocks is the CTAs-per-SM bound of the instantiation (chosen by the host from the shared-memory footprint),
// i.e. the register budget: 5 -> <=100 registers, 6 -> <=80, 7 -> <=72, 8 -> 64, 10 -> <=51.  CV is the config view.
template <int kThreads, int kMinBlocks, class CV>
__global__ void __launch_bounds__(kThreads, kMinBlocks) agar_step_kernel(const __grid_constant__ DevParams p,
                                                               const __grid_constant__ KArgs a) {
    extern __shared__ __align__(16) float smem_raw[];
    decltype(auto) L = layout_of<CV>(p);
    const AgarConfig& c = p.c;
    const int G = CV_INT(G, c.grid_size), P = CV_INT(P, c.num_pellets), TICKS = CV_INT(TICKS, c.ticks_per_step);
    const Smem s = carve(smem_raw, L, G);
    const int tid = threadIdx.x, nt = kThreads, lane = tid & 31;
    const int hbase = lane & 16, slot = lane & 15;
    int arena = a.arena0 + blockIdx.x;
    if (a.order_in) {  // the lists of the previous launch, costliest class first
        const int4 cnt = *reinterpret_cast<const int4*>(a.order_in);
        int b = blockIdx.x, cls = 0;
        if (b >= cnt.x) { b -= cnt.x; cls = 1; if (b >= cnt.y) { b -= cnt.y; cls = 2; if (b >= cnt.z) { b -= cnt.z; cls = 3; } } }
        arena = a.order_in[8 + cls * a.order_stride + b];
    }
    const long long t_begin = (a.order_out && tid == 0) ? clock64() : 0;
    if (a.order_clear && blockIdx.x == 0 && tid == 0) {
        *reinterpret_cast<int4*>(a.order_clear) = make_int4(0, 0, 0, 0);
        *reinterpret_cast<int4*>(a.order_clear + 4) = make_int4(0, 0, 0, 0);
    }
    const int A = CV_INT(A, c.num_agents), K = CV_INT(K, p.K), KC = CV_INT(KC, p.KC), V = CV_INT(V, c.num_viruses), F = CV_INT(F, c.max_foods);
    const uint64_t gid64 = (uint64_t)(c.arena_id_base + arena);
    const uint32_t gid_lo = (uint32_t)gid64, gid_hi = (uint32_t)(gid64 >> 32);
    const float r2pm = c.r2_per_mass, arena_size = c.arena_size;
    float* cx = s.rec + L.cell_x; float* cy = s.rec + L.cell_y; float* cix = s.rec + L.cell_ix;
    float* ciy = s.rec + L.cell_iy; float* cm = s.rec + L.cell_m; float* ct = s.rec + L.cell_t;
    float* px = s.rec + L.pellet_x; float* py = s.rec + L.pellet_y;
    float* vx = s.rec + L.virus_x; float* vy = s.rec + L.virus_y;
    float* fx = s.rec + L.food_x; float* fy = s.rec + L.food_y; float* fvx = s.rec + L.food_vx; float* fvy = s.rec + L.food_vy;
    float* grec = a.state + (size_t)arena * L.words;  // write target
    int32_t* inl = s.sc + 32;                         // bitmask of gids present in the alive list (sc[0..31] are scalars)

    PHASE_DECL
    // ---- stage the record (and the persisted pellet hash) with TMA bulk copies ----
    const int hash_words = CV::is_static ? (int)hash_record_words(L) : p.hash_words;
    int32_t* ghash = a.hash ? a.hash + (size_t)arena * hash_words : nullptr;
    const bool load_hash = (a.flags & F_STEP) && ghash;
    if (tid == 0) mbar_init(s.mbar, 1);
    __syncthreads();
    if (tid == 0) {
        const uint32_t rec_bytes = (uint32_t)L.words * 4u, hash_bytes = load_hash ? (uint32_t)hash_words * 4u : 0u;
        mbar_expect_tx(s.mbar, rec_bytes + hash_bytes);
        bulk_g2s(s.rec, a.src_state + (size_t)arena * L.words, rec_bytes, s.mbar);
        if (load_hash) bulk_g2s(s.hstart, ghash, hash_bytes, s.mbar);
    }
    {   // meanwhile: clear the per-tick scratch
        // (64-bit stores: every scratch array starts on an 8-byte boundary and holds an even number of words)
        const uint2 ones2 = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu), zero2 = make_uint2(0u, 0u);
        for (int i = tid; i < KC / 4; i += nt) {
            reinterpret_cast<uint2*>(s.cnt32)[i] = zero2;
            reinterpret_cast<uint2*>(s.eater)[i] = ones2;
        }
        for (int i = tid; i < (int)gain_words(L) / 2; i += nt) reinterpret_cast<uint2*>(s.gain)[i] = zero2;
        if (tid < 64) s.sc[tid] = 0;
        if (a.flags & F_STEP) {
            for (int i = tid; i < L.P_pad / 4; i += nt) reinterpret_cast<uint2*>(s.pwin)[i] = ones2;
        }
    }
    // the agents' actions (two dependent global loads each) are fetched while the record is in flight
    float act_x = 0.0f, act_y = 0.0f; int act_a = 0;
    if ((a.flags & F_STEP) && tid < A) {
        const size_t ia = (size_t)arena * A + tid;
        if (a.act_index) {
            int idx = a.act_index[ia];
            idx = idx < 0 ? 0 : (idx >= a.n_actions ? a.n_actions - 1 : idx);
            act_x = a.tab_xy[2 * idx]; act_y = a.tab_xy[2 * idx + 1]; act_a = a.tab_act[idx];
        } else {
            act_x = a.target_xy[2 * ia]; act_y = a.target_xy[2 * ia + 1]; act_a = a.act[ia];
        }
    }
    mbar_wait(s.mbar, 0);
    __syncthreads();
    // persisted pellet hash of this arena: valid if it carries the handle's current epoch
    int4 hhdr = make_int4(0, 0, 0, 0);
    if (load_hash) hhdr = *reinterpret_cast<const int4*>(s.hhdr);

    bool static_dirty = false;  // whole record must be written back (after a reset)
    int persist_hash = 0;       // 0: the persisted hash is current, 1: its slot order changed, 2: rebuilt
    if (a.flags & F_RESET) {
        if (!a.reset_mask || a.reset_mask[arena]) {
            reset_arena<CV>(p, gid_lo, gid_hi); static_dirty = true;
            // the pellets of this arena (only) were redrawn: its persisted hash is stale
            if (tid == 0 && a.hash && !(a.flags & F_READONLY))
                *reinterpret_cast<int4*>(a.hash + (size_t)arena * hash_words + hash_body_words(L)) = make_int4(0, 0, 0, 0);
        }
    }

    if (a.flags & F_STEP) {
        const uint32_t tick0 = (uint32_t)s.reci[L.tick];
        PHASE(0);  // staging
        // ---- pellet hash (persisted across steps, rebuilt when stale), alive list ----
        if (load_hash && hhdr.x == p.hash_epoch && !static_dirty) {
            if (tid == 0) s.sc[SC_HASHOK] = 1;
        } else {
            build_hash<CV>(p);
        }
        build_list<CV>(p, s);
        __syncthreads();
        PHASE(1);  // hash check/build, list
        // ---- players: bot respawn, centroid and mass of single-cell players (thread per player) ----
        for (int k = tid; k < K; k += nt) {
            unsigned pm = (unsigned)s.pmask[k];
            if (k >= A && pm == 0u) {  // dead bot: respawn (spawn_player in the oracle)
                const int g = k * MC;
                float x, y;
                rand_pos(p, gid_lo, gid_hi, (uint32_t)k, AGAR_STREAM_PLAYER_SPAWN, tick0, x, y);
                cx[g] = x; cy[g] = y; cm[g] = c.start_mass;  // the other fields of a dead player's slots are 0
                ev_push<CV>(p, a, s, arena, 0, AGAR_EV_RESPAWN, k, 0);
                pm = 1u; s.pmask[k] = 1;
                if (!(atomicOr(&inl[g >> 5], 1 << (g & 31)) & (1 << (g & 31)))) s.list[atomicAdd(&s.sc[SC_NLIST], 1)] = (uint16_t)g;
            }
            s.palive[k] = pm != 0u;
            s.Tx[k] = 0.0f; s.Ty[k] = 0.0f;
            if (k < A) { s.pact[k] = 0; s.dirx[k] = 1.0f; s.diry[k] = 0.0f; }
            float Cx = 0.0f, Cy = 0.0f, mb = 0.0f;
            if (__popc(pm) == 1) {  // np.average over one cell: (x*m)/m, exactly as the general path gives
                const int g = k * MC + __ffs(pm) - 1;
                const float m = cm[g];
                Cx = (cx[g] * m) / m; Cy = (cy[g] * m) / m; mb = m;
            }
            s.Cx[k] = Cx; s.Cy[k] = Cy;
            if (k < A) s.mass0[k] = mb;
        }
        __syncthreads();
        // ---- players with several cells: numpy-order centroid and butterfly mass (half-warp per player) ----
        if (s.sc[SC_MULTI]) {
            for (int base = (tid & ~31); base < KC; base += nt) {
                const int g = base + lane, k = g >> 4;
                const bool multi = g < KC && __popc((unsigned)s.pmask[k]) >= 2;
                if (!__any_sync(FULL, multi)) continue;  // warp-uniform
                const float m = multi ? cm[g] : 0.0f;
                float Cx = 0.0f, Cy = 0.0f;
                hw_centroid(hbase, slot, multi ? cx[g] : 0.0f, multi ? cy[g] : 0.0f, m, Cx, Cy);
                const float mb = hw_treesum(m);
                if (multi && slot == 0) { s.Cx[k] = Cx; s.Cy[k] = Cy; if (k < A) s.mass0[k] = mb; }
            }
            __syncthreads();
        }
        // ---- agents: action -> target point, direction, act ----
        if (tid < A && s.palive[tid]) {  // A <= 64 <= nt: thread k is agent k (its action was loaded during staging)
            const int k = tid;
            float ax = act_x, ay = act_y; int act = act_a;
            float n2 = ax * ax + ay * ay;
            if (!(n2 < INFINITY)) { ax = 0.0f; ay = 0.0f; }
            else if (n2 > 1.0f) { float inv = 1.0f / sqrtf(n2); ax = ax * inv; ay = ay * inv; }
            s.Tx[k] = s.Cx[k] + ax * c.target_reach;
            s.Ty[k] = s.Cy[k] + ay * c.target_reach;
            float m2 = ax * ax + ay * ay;
            if (m2 > 0.0f) { float inv = 1.0f / sqrtf(m2); s.dirx[k] = ax * inv; s.diry[k] = ay * inv; }
            act = (act == AGAR_ACT_SPLIT || act == AGAR_ACT_FEED) ? act : 0;
            s.pact[k] = act;
            if (act == AGAR_ACT_FEED) s.sc[SC_ANYFEED] = 1;
            if (act == AGAR_ACT_SPLIT) s.sc[SC_ANYSPLIT] = 1;
        }
        PHASE(2);  // players, multi, agent actions
        // ---- bots: nearest pellet in the square sense window, else the roam way-point.  Four lanes
        //      per bot search the 3x3 bins around it; the result is final when it is closer than the
        //      nearest unsearched bin edge, otherwise the whole window is searched ----
        {
            const int l4 = lane & 3;
            for (int kb = A + (tid >> 5) * 8; kb < K; kb += nt >> 2) {  // 8 bots per warp pass, warp-uniform
                const int k = min(kb + (lane >> 2), K - 1);
                const bool valid = kb + (lane >> 2) < K;
                const float Cx = s.Cx[k], Cy = s.Cy[k], R = c.bot_sense_radius;
                const int bx = hash_coord(p, Cx), by = hash_coord(p, Cy);
                int x0 = max(bx - 1, 0), x1 = min(bx + 1, HASH_W - 1), y0 = max(by - 1, 0), y1 = min(by + 1, HASH_W - 1);
                float best = INFINITY; int bi = NONE_I;
                for (int pass = 0; pass < 2; ++pass) {
                    for (int row = y0; row <= y1; ++row) {
                        const int j1 = s.hstart[row * HASH_W + x1 + 1];
                        for (int j = s.hstart[row * HASH_W + x0] + l4; j < j1; j += 4) {
                            const int i = s.sorted[j];
                            if (i == (int)HASH_FREE) continue;
                            float dx = px[i] - Cx, dy = py[i] - Cy;
                            if (fabsf(dx) <= R && fabsf(dy) <= R) {
                                float d2 = dx * dx + dy * dy;
                                if (d2 < best || (d2 == best && i < bi)) { best = d2; bi = i; }
                            }
                        }
                    }
#pragma unroll
                    for (int o = 2; o > 0; o >>= 1) {
                        float ob = __shfl_xor_sync(FULL, best, o);
                        int oi = __shfl_xor_sync(FULL, bi, o);
                        if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
                    }
                    if (pass == 1) break;  // warp-uniform: `pass` only advances together (see below)
                    // distance from the bot to the nearest edge of the searched block that has bins beyond it
                    float margin = INFINITY;
                    if (x0 > 0) margin = fminf(margin, Cx - (float)x0 * p.hash_bs);
                    if (x1 < HASH_W - 1) margin = fminf(margin, (float)(x1 + 1) * p.hash_bs - Cx);
                    if (y0 > 0) margin = fminf(margin, Cy - (float)y0 * p.hash_bs);
                    if (y1 < HASH_W - 1) margin = fminf(margin, (float)(y1 + 1) * p.hash_bs - Cy);
                    const float safe = margin * 0.999f - 0.01f;
                    const bool done = !valid || (bi != NONE_I && safe > 0.0f && best < safe * safe) || safe >= R * 1.5f;
                    if (__all_sync(FULL, done)) break;  // warp-uniform exit
                    if (done) { x0 = 1; x1 = 0; y0 = 1; y1 = 0; continue; }  // this quad keeps its result: empty second pass
                    const float Rp = R * 1.001f + 0.01f;
                    x0 = hash_coord(p, Cx - Rp); x1 = hash_coord(p, Cx + Rp);
                    y0 = hash_coord(p, Cy - Rp); y1 = hash_coord(p, Cy + Rp);
                    best = INFINITY; bi = NONE_I;
                }
                if (valid && l4 == 0) {
                    if (bi != NONE_I) { s.Tx[k] = px[bi]; s.Ty[k] = py[bi]; }
                    else {
                        uint32_t epoch = tick0 & ~(uint32_t)(c.bot_roam_period - 1);
                        float x, y;
                        rand_pos(p, gid_lo, gid_hi, (uint32_t)k, AGAR_STREAM_BOT_ROAM, epoch, x, y);
                        s.Tx[k] = x; s.Ty[k] = y;
                    }
                }
            }
        }
        __syncthreads();
        PHASE(3);  // bots
        // ---- split: i-th eligible cell -> i-th free slot (ballot ranks inside the half-warp) ----
        if (s.sc[SC_ANYSPLIT]) {
            for (int base = (tid & ~31); base < A * MC; base += nt) {
                const int g = base + lane, k = g >> 4;
                const bool splitting = g < A * MC && s.pact[k] == AGAR_ACT_SPLIT;
                if (!__any_sync(FULL, splitting)) continue;  // warp-uniform
                float m = splitting ? cm[g] : -1.0f;
                unsigned freeb = (__ballot_sync(FULL, m == 0.0f) >> hbase) & 0xFFFFu;
                bool elig = splitting && m >= c.split_min_mass;
                unsigned eligb = (__ballot_sync(FULL, elig) >> hbase) & 0xFFFFu;
                int i = __popc(eligb & ((1u << slot) - 1u));
                if (elig && i < __popc(freeb)) {
                    int d = k * MC + (int)__fns(freeb, 0, i + 1);
                    float halfm = m * 0.5f;
                    cm[g] = halfm; ct[g] = c.merge_delay;
                    cx[d] = cx[g]; cy[d] = cy[g];
                    cix[d] = s.dirx[k] * c.split_speed; ciy[d] = s.diry[k] * c.split_speed;
                    cm[d] = halfm; ct[d] = c.merge_delay;
                    ev_push<CV>(p, a, s, arena, 0, AGAR_EV_SPLIT, g, d);
                    s.sc[SC_MULTI] = 1;
                    if (!(c.merge_delay > 0.0f)) s.sc[SC_MERGEABLE] = 1;
                    if (!(atomicOr(&inl[d >> 5], 1 << (d & 31)) & (1 << (d & 31)))) s.list[atomicAdd(&s.sc[SC_NLIST], 1)] = (uint16_t)d;
                }
            }
        }
        // ---- feed: cells in gid order take free food slots in ascending order (warp 0) ----
        if (s.sc[SC_ANYFEED] && tid < 32) {
            int32_t* freefood = reinterpret_cast<int32_t*>(s.gain);  // gain[] rests at 0 outside the eat phase: borrowed
            int nfree = 0;
            for (int base = 0; base < L.F_pad; base += 32) {
                int f = base + lane;
                bool fr = f < F && !(fx[f] >= 0.0f);
                unsigned b = __ballot_sync(0xffffffffu, fr);
                if (fr) freefood[nfree + __popc(b & ((1u << lane) - 1u))] = f;
                nfree += __popc(b);
            }
            __syncwarp();
            int used = 0;
            for (int k = 0; k < A; ++k) {
                if (s.pact[k] != AGAR_ACT_FEED) continue;
                int g = k * MC + lane;
                float m = (lane < MC) ? cm[g] : 0.0f;
                bool elig = (lane < MC) && m >= c.feed_min_mass;
                unsigned b = __ballot_sync(0xffffffffu, elig);
                int rank = used + __popc(b & ((1u << lane) - 1u));
                if (elig && rank < nfree) {
                    int f = freefood[rank];
                    float r = sqrtf(m * r2pm);
                    cm[g] = m - c.food_mass;
                    float off = r + c.food_spawn_gap;
                    fx[f] = clampf(cx[g] + s.dirx[k] * off, 0.0f, arena_size);
                    fy[f] = clampf(cy[g] + s.diry[k] * off, 0.0f, arena_size);
                    fvx[f] = s.dirx[k] * c.feed_speed; fvy[f] = s.diry[k] * c.feed_speed;
                    ev_push<CV>(p, a, s, arena, 0, AGAR_EV_FEED, g, f);
                }
                used += __popc(b);
            }
            __syncwarp();
            for (int e = lane; e < nfree; e += 32) s.gain[e] = 0.0f;  // hand gain[] back zeroed
        }
        __syncthreads();

        PHASE(4);  // split, feed
        for (int tk = 0; tk < TICKS; ++tk) {
            const uint32_t tick = tick0 + (uint32_t)tk;
            if (s.sc[SC_REBUILD]) build_hash<CV>(p);
            PHASE(5);  // in-tick rebuild  // too many respawns since the last build (uniform)
            // ---- motion + pellet query: four lanes per alive cell.  All four evaluate the motion (lane 0
            //      stores it) and then share the scan of the hash runs under the cell's padded bounding
            //      box and of the respawn list.  A pellet whose centre lies inside the disc is claimed
            //      with atomicMin: the lowest gid wins ----
            {
                const int n = s.sc[SC_NLIST];
                const int l4 = lane & 3;
                for (int qb = (tid >> 5) * 8; qb < n; qb += nt >> 2) {  // 8 cells per warp pass, warp-uniform
                    const int q = qb + (lane >> 2);
                    const int g = q < n ? s.list[q] : s.list[0];
                    float m = q < n ? cm[g] : 0.0f;
                    const float ox = cx[g], oy = cy[g], oix = cix[g], oiy = ciy[g], t = ct[g];
                    __syncwarp();  // lane 0 of a quad stores the cell below: its whole quad has read it by now
                    if (!(m > 0.0f)) continue;  // no cell, or it died earlier in this step
                    if (c.mass_decay > 0.0f) {  // mass decay (own-spec rule, off by default): never below start_mass
                        const float md = m * (1.0f - c.mass_decay);
                        if (md >= c.start_mass) { m = md; if (l4 == 0) cm[g] = m; }
                    }
                    const int k = g >> 4;
                    const float r2 = m * r2pm;
                    const float r = sqrtf(r2);
                    float spd = c.speed_k / sqrtf(r);
                    float dx = s.Tx[k] - ox, dy = s.Ty[k] - oy;
                    float d = sqrtf(dx * dx + dy * dy);
                    float den = d > c.target_reach ? d : c.target_reach;
                    float sc = spd / den;
                    float nx = ox + (dx * sc + oix);
                    float ny = oy + (dy * sc + oiy);
                    float ix = oix * c.impulse_decay, iy = oiy * c.impulse_decay;
                    if (fabsf(ix) < 1e-3f && fabsf(iy) < 1e-3f) { ix = 0.0f; iy = 0.0f; }
                    const float x = clampf(nx, 0.0f, arena_size), y = clampf(ny, 0.0f, arena_size);
                    if (l4 == 0) {
                        cx[g] = x; cy[g] = y; cix[g] = ix; ciy[g] = iy;
                        if (t > 0.0f) {
                            ct[g] = t - 1.0f;
                            if (t - 1.0f == 0.0f) s.sc[SC_MERGEABLE] = 1;  // a merge becomes possible (conservative: set per cell)
                        }
                    }
                    const float rp = r * 1.001f + 0.01f;
                    const int bx0 = hash_coord(p, x - rp), bx1 = hash_coord(p, x + rp);
                    const int by0 = hash_coord(p, y - rp), by1 = hash_coord(p, y + rp);
                    for (int by = by0; by <= by1; ++by) {
                        const int j1 = s.hstart[by * HASH_W + bx1 + 1];
                        for (int j = s.hstart[by * HASH_W + bx0] + l4; j < j1; j += 4) {
                            const int i = s.sorted[j];
                            if (i == (int)HASH_FREE) continue;
                            const float ex = px[i] - x, ey = py[i] - y;
                            if (ex * ex + ey * ey < r2)
                                if (claim_min16(&s.pwin[i], g)) { const int e = atomicAdd(&s.sc[SC_NHITS], 1); if (e < HITS_CAP) s.hits[e] = (uint16_t)i; }
                        }
                    }
                }
#pragma unroll 1
                for (int f = tid; f < F; f += nt)
                    if (fx[f] >= 0.0f) s.sc[SC_FOODANY] = 1;
                for (int v = tid; v < V; v += nt) s.vwin[v] = NONE_I;
            }
            __syncthreads();
            PHASE(6);  // motion + query (+ barrier)
            const int nh = s.sc[SC_NHITS];                // block-uniform from here on
            const bool food_any = s.sc[SC_FOODANY] != 0;
            const bool gains = nh > 0 || food_any;        // some mass changes this tick before the eat test
            // ---- resolve the claimed pellets: winner = lowest gid (atomicMin), respawn by Philox ----
            if (nh > 0) {
                const bool listed = nh <= HITS_CAP;  // else the list overflowed: every pellet's winner slot is examined
                const int nscan = listed ? nh : P;
                for (int h = tid; h < nscan; h += nt) {
                    const int i = listed ? (int)s.hits[h] : h;
                    const int g = s.pwin[i];
                    if (g == NONE16) continue;
                    s.pwin[i] = NONE16;
                    cnt_inc(s.cnt32, g);
                    ev_push<CV>(p, a, s, arena, tk, AGAR_EV_PELLET_EATEN, g, i);
                    float nx = -1.0f, ny = -1.0f;
                    const bool hash_live = s.sc[SC_HASHOK] != 0;  // (it always is while ticks run; kept for safety)
                    if (hash_live) hash_remove(p, s, i, px[i], py[i]);
                    if (c.pellet_regen) {
                        rand_pos(p, gid_lo, gid_hi, (uint32_t)i, AGAR_STREAM_PELLET_RESPAWN, tick, nx, ny);
                        if (hash_live && !hash_insert(p, s, i, nx, ny)) s.sc[SC_REBUILD] = 1;  // its bin is full: rebuild next tick
                    }
                    s.sc[SC_HDIRTY] = 1;
                    px[i] = nx; py[i] = ny;
                    if (!(a.flags & F_READONLY)) { grec[L.pellet_x + i] = nx; grec[L.pellet_y + i] = ny; }
                }
                __syncthreads();
            }
            if (food_any) {
                // ---- pellet gains, then foods: move, eaten by the lowest-gid cell holding the centre ----
                const int n = s.sc[SC_NLIST];
                for (int q = tid; q < n; q += nt) {
                    const int g = s.list[q];
                    const int e = cnt_get(s.cnt32, g);
                    if (e) { cm[g] = cm[g] + (float)e * c.pellet_mass; cnt_clear(s.cnt32, g); }
                }
                __syncthreads();
                // a lane moves its food; then the warp scans the alive cells for every moved food in turn (foods are few)
                for (int base = 0; base < F; base += nt) {  // warp-uniform; slots interleaved over the warps
                    const int f = base + lane * (nt >> 5) + (tid >> 5);
                    float x = f < F ? fx[f] : -1.0f, y = 0.0f;
                    const bool act = x >= 0.0f;
                    if (act) {
                        const float nx = x + fvx[f], ny = fy[f] + fvy[f];
                        float ux = fvx[f] * c.food_decay, uy = fvy[f] * c.food_decay;
                        if (fabsf(ux) < 1e-3f && fabsf(uy) < 1e-3f) { ux = 0.0f; uy = 0.0f; }
                        x = clampf(nx, 0.0f, arena_size); y = clampf(ny, 0.0f, arena_size);
                        fx[f] = x; fy[f] = y; fvx[f] = ux; fvy[f] = uy;
                    }
                    for (unsigned todo = __ballot_sync(FULL, act); todo; todo &= todo - 1) {
                        const int src = __ffs(todo) - 1;
                        const float xs = __shfl_sync(FULL, x, src), ys = __shfl_sync(FULL, y, src);
                        int best = NONE_I;
                        for (int q = lane; q < n; q += 32) {
                            const int g = s.list[q];
                            const float m = cm[g];
                            const float dx = xs - cx[g], dy = ys - cy[g];
                            if (m > 0.0f && dx * dx + dy * dy < m * r2pm && g < best) best = g;
                        }
                        best = __reduce_min_sync(FULL, best);
                        if (lane == src && best != NONE_I) {
                            cnt_inc(s.cnt32, best);
                            ev_push<CV>(p, a, s, arena, tk, AGAR_EV_FOOD_EATEN, best, f);
                            fx[f] = -1.0f; fy[f] = 0.0f; fvx[f] = 0.0f; fvy[f] = 0.0f;
                        }
                    }
                }
                __syncthreads();
            }
            PHASE(7);  // resolve, foods
            // cell-eats-cell test (different players, pre-phase snapshot): four lanes per prey, each
            // scanning a quarter of the possible eaters; the lowest gid wins
            // Measured and dropped (profiles/r01i_notes.txt): culling the scan through per-bin cell counts or bin
            // heads shortens a lone CTA's critical path but not the throughput at 7 CTAs per SM.
            // With several players the same quads also make the heavy cells' virus claims (first pass only), each lane
            // testing a quarter of the viruses.
            auto eat_test = [&](bool claim_viruses) {
                const int n = s.sc[SC_NLIST];
                const int l4 = lane & 3;
                for (int qb = (tid >> 5) * 8; qb < n; qb += nt >> 2) {  // 8 preys per warp pass, warp-uniform
                    const int q = qb + (lane >> 2);
                    const int j = q < n ? s.list[q] : s.list[0];
                    const float mj = q < n ? cm[j] : 0.0f;
                    const float need = mj * c.eat_ratio, x = cx[j], y = cy[j];
                    const int kj = j >> 4;
                    int best = NONE_I;
                    if (mj > 0.0f) {
#pragma unroll 1  // measured: 1 -> 29.7 M, 2 -> 29.6 M, 4 -> 29.3 M, 8 -> 29.0 M env-steps/s on cfg2 (smaller code wins)
                        for (int qi = l4; qi < n; qi += 4) {
                            const int i = s.list[qi];
                            const float mi = cm[i];
                            if ((i >> 4) == kj || !(mi >= need)) continue;  // need > 0, so dead cells fail here
                            const float dx = x - cx[i], dy = y - cy[i];
                            if (dx * dx + dy * dy < mi * r2pm && i < best) best = i;
                        }
                    }
                    best = min(best, __shfl_xor_sync(FULL, best, 1));
                    best = min(best, __shfl_xor_sync(FULL, best, 2));
                    if (l4 == 0 && best != NONE_I) { s.eater[j] = (uint16_t)best; s.sc[SC_ANYEAT] = 1; }
                    if (claim_viruses && mj >= c.virus_pop_mass) {  // lowest-gid heavy cell holding a virus centre wins it
                        const float r2 = mj * r2pm;
                        for (int v = l4; v < V; v += 4) {
                            float dx = vx[v] - x, dy = vy[v] - y;
                            if (dx * dx + dy * dy < r2) { atomicMin(&s.vwin[v], j); s.sc[SC_ANYVHIT] = 1; }
                        }
                    }
                }
            };
            // ---- apply gains (pellets, or foods when present); in a one-player world heavy cells claim viruses here
            //      (else the eat-test quads do).  On a tick without gains no mass changes here, so the eat test runs
            //      in the same phase ----
            if (gains || (V > 0 && K == 1)) {
                const int n = s.sc[SC_NLIST];
                const float unit = food_any ? c.food_mass : c.pellet_mass;
                for (int q = tid; q < n; q += nt) {
                    const int g = s.list[q];
                    float m = cm[g];
                    const int e = cnt_get(s.cnt32, g);
                    if (e) { m = m + (float)e * unit; cm[g] = m; cnt_clear(s.cnt32, g); }
                    if (V > 0 && K == 1 && m >= c.virus_pop_mass) {
                        const float r2 = m * r2pm, x = cx[g], y = cy[g];
                        for (int v = 0; v < V; ++v) {
                            float dx = vx[v] - x, dy = vy[v] - y;
                            if (dx * dx + dy * dy < r2) { atomicMin(&s.vwin[v], g); s.sc[SC_ANYVHIT] = 1; }
                        }
                    }
                }
            }
            if (tid == 0 && gains) { s.sc[SC_NHITS] = 0; s.sc[SC_FOODANY] = 0; }  // a barrier follows whenever gains
            if (gains && K > 1) __syncthreads();  // the eat test reads the updated masses
            PHASE(8);  // apply gains, virus claims
            // ---- cell eats cell (one code site).  It runs speculatively before the virus pops are known; a
            //      pop (rare) changes masses, voids the result and the test is taken again ----
            for (bool again = true, first = true; again; first = false) {
                if (K > 1) eat_test(first && V > 0);
                if (gains || V > 0 || K > 1) __syncthreads();
                again = false;
                if (V > 0 && s.sc[SC_ANYVHIT]) {  // block-uniform, rare: out of line
                    virus_pops<CV>(p, a, arena, tk, tick, gid_lo, gid_hi);
                    again = K > 1;
                }
            }
            PHASE(9);  // eat test (+ pops) + barrier
            if (K > 1) {
                if (s.sc[SC_ANYEAT]) {
                    const int n = s.sc[SC_NLIST];
                    if (tid < 32) {  // gains added in ascending prey order (float sum order is part of the spec)
                        for (int base = 0; base < KC; base += 32) {
                            int j = base + lane;
                            int w = (j < KC) ? (int)s.eater[j] : NONE16;
                            unsigned mask = __ballot_sync(0xffffffffu, w != NONE16);
                            while (mask) {
                                int src = __ffs(mask) - 1;
                                if (lane == src) {
                                    s.gain[w] = s.gain[w] + cm[j];
                                    ev_push<CV>(p, a, s, arena, tk, AGAR_EV_CELL_EATEN, w, j);
                                }
                                __syncwarp();
                                mask &= mask - 1;
                            }
                        }
                    }
                    __syncthreads();
                    for (int q = tid; q < n; q += nt) {
                        const int g = s.list[q];
                        if (s.eater[g] != NONE16) { clear_cell<CV>(p, s, g); s.eater[g] = NONE16; }  // eaten: forfeits its gains
                        else if (s.gain[g] > 0.0f) cm[g] = cm[g] + s.gain[g];
                        s.gain[g] = 0.0f;
                    }
                    if (tid == 0) s.sc[SC_ANYEAT] = 0;
                    __syncthreads();
                }
            }
            PHASE(10);  // eat apply
            // ---- merge own cells whose timers ran out (half-warp per player, shuffles).  Skipped unless some player
            //      held two cells with run-out timers at the start of the step or a timer ran out since ----
            if (s.sc[SC_MULTI] && s.sc[SC_MERGEABLE]) {
                for (int base = (tid & ~31); base < KC; base += nt) {
                    const int g = base + lane;
                    const bool valid = g < KC;
                    const float m = valid ? cm[g] : 0.0f, t = valid ? ct[g] : 1.0f, x = valid ? cx[g] : 0.0f, y = valid ? cy[g] : 0.0f;
                    const bool can = m > 0.0f && t == 0.0f;
                    unsigned canb = (__ballot_sync(FULL, can) >> hbase) & 0xFFFFu;
                    if (!__any_sync(FULL, __popc(canb) >= 2)) continue;  // warp-uniform
                    const float r2 = m * r2pm;
                    int root = slot;
#pragma unroll 1
                    for (int q = 0; q < MC - 1; ++q) {
                        float xa = __shfl_sync(FULL, x, hbase + q), ya = __shfl_sync(FULL, y, hbase + q);
                        float ra = __shfl_sync(FULL, r2, hbase + q);
                        if (can && root == slot && q < slot && ((canb >> q) & 1u)) {
                            float dx = x - xa, dy = y - ya;
                            float rr = ra > r2 ? ra : r2;
                            if (dx * dx + dy * dy < rr) root = q;  // partner = lowest qualifying slot below
                        }
                    }
                    // follow partners down to the final survivor (partner < slot, so 4 jumps suffice)
#pragma unroll
                    for (int it = 0; it < 4; ++it) root = __shfl_sync(FULL, root, hbase + root);
                    unsigned merged = (__ballot_sync(FULL, root != slot) >> hbase) & 0xFFFFu;
                    if (!__any_sync(FULL, merged != 0u)) continue;  // warp-uniform
                    float acc = m;
#pragma unroll 1
                    for (int q = 1; q < MC; ++q) {
                        float mq = __shfl_sync(FULL, m, hbase + q);
                        int rq = __shfl_sync(FULL, root, hbase + q);
                        if (rq == slot && q != slot) acc = acc + mq;
                    }
                    if (root != slot) {
                        ev_push<CV>(p, a, s, arena, tk, AGAR_EV_MERGE, (g & ~15) + root, g);
                        clear_cell<CV>(p, s, g);
                    } else if (acc != m) {
                        cm[g] = acc;
                    }
                }
                __syncthreads();
            }
        }

        PHASE(11);  // merge (last tick only; earlier ticks charge it to the rebuild slot)
        // ---- rewards / dones ----
        for (int base = (tid & ~31); base < A * MC; base += nt) {
            const int g = base + lane, k = g >> 4;
            const bool valid = g < A * MC;
            float m = valid ? cm[g] : 0.0f;
            float after = hw_treesum(m);
            bool alive = ((__ballot_sync(FULL, m > 0.0f) >> hbase) & 0xFFFFu) != 0u;
            if (valid && slot == 0) {
                int was = s.reci[L.agent_done + k];
                float rw = was ? 0.0f : after - s.mass0[k];
                if (!was && !alive) {
                    s.reci[L.agent_done + k] = 1;
                    ev_push<CV>(p, a, s, arena, TICKS, AGAR_EV_DEATH, k, 0);
                }
                int d = s.reci[L.agent_done + k];
                a.reward[(size_t)arena * A + k] = rw;
                a.done[(size_t)arena * A + k] = (uint8_t)d;
                if (!d) s.sc[SC_ALLDONE] = 1;  // at least one agent still playing
            }
        }
        if (tid == 0) {
            s.reci[L.tick] = (int32_t)(tick0 + (uint32_t)TICKS);
            s.reci[L.episode_steps] += 1;
        }
        __syncthreads();
        if (c.auto_reset && !s.sc[SC_ALLDONE]) {
            reset_arena<CV>(p, gid_lo, gid_hi); static_dirty = true;
            if (tid == 0) s.sc[SC_HASHOK] = 0;  // pellets were redrawn: the raster scans them all
        }
        PHASE(12);  // rewards, auto reset
        // ---- persist the pellet hash for the next step (the copy itself goes out with the record below) ----
        if (a.hash && !(a.flags & F_READONLY)) {
            const bool ok = s.sc[SC_HASHOK] != 0 && s.sc[SC_REBUILD] == 0;  // else redrawn pellets / a full bin: rebuild next step
            persist_hash = !ok ? 0 : (s.sc[SC_HBUILT] ? 2 : (s.sc[SC_HDIRTY] ? 1 : 0));  // 2: bins + order, 1: order only
            if (tid == 0 && (!ok || s.sc[SC_HBUILT] || hhdr.x != p.hash_epoch))
                *reinterpret_cast<int4*>(a.hash + (size_t)arena * hash_words + hash_body_words(L)) = make_int4(ok ? p.hash_epoch : 0, 0, 0, 0);
        }
        if (!CV::is_static && c.event_capacity > 0 && tid == 0) a.ev_counts[arena] = s.sc[SC_EVCOUNT];
    }

    PHASE(13);  // persist hash
    // ---- write back: one bulk store of the rewritten part (the whole record after a reset) ----
    if (!(a.flags & F_READONLY) && (a.flags & (F_STEP | F_RESET))) {
        const int w0 = static_dirty ? 0 : L.dyn_begin;  // dyn_begin % 4 == 0
        __syncthreads();                 // every thread's writes to the record are done
        if (tid == 0) {
            // ... and ordered before the async proxy's reads of shared memory (the narrow fence: the full one also
            // waits for this thread's global stores).  After a reset the bulk store covers the pellet / virus words
            // that respawns wrote to global memory with plain stores earlier in this step: only then must the global
            // side be ordered too.
            if (static_dirty) fence_async_all(); else fence_async_smem();
            bulk_s2g(grec + w0, s.rec + w0, (uint32_t)(L.words - w0) * 4u);
            if (persist_hash) {          // [hstart | sorted] or only the slots; the raster below only reads them
                const int h0 = persist_hash == 2 ? 0 : HASH_BINS + 4;
                bulk_s2g(a.hash + (size_t)arena * hash_words + h0, s.hstart + h0, (uint32_t)((int)hash_body_words(L) - h0) * 4u);
            }
            bulk_commit();
        }
    }
    PHASE(14);  // write back
    // ---- file this arena for the next launch: class by its duration up to here (the raster costs every arena the same).
    //      The slot comes back from the atomic while the raster runs; the entry is stored after it ----
    int order_slot = -1, order_cls = 3;
    if (a.order_out && tid == 0) {
        const unsigned long long cyc = (unsigned long long)(clock64() - t_begin);
        if (a.order_in) {
            const unsigned long long mean = *reinterpret_cast<const unsigned long long*>(a.order_in + 4) / (unsigned long long)a.n_records;
            order_cls = cyc > 2 * mean ? 0 : (cyc > mean + (mean >> 2) ? 1 : (cyc > mean - (mean >> 3) ? 2 : 3));
        }
        order_slot = atomicAdd(&a.order_out[order_cls], 1);
        atomicAdd(reinterpret_cast<unsigned long long*>(a.order_out + 4), cyc);
    }
    // ---- observations ----
    if ((a.flags & F_OBS) && a.obs) {
        __syncthreads();
        const int GG = G * G;
        raster_arena<CV>(p, s, a.obs + (size_t)arena * A * CV_INT(C, p.C) * GG, a.obs_meta ? a.obs_meta + (size_t)arena * A * 4 : nullptr,
                         s.sc[SC_HASHOK] != 0 && s.sc[SC_REBUILD] == 0,
                         ((a.flags & F_STEP) && !static_dirty && KC > 256) ? s.sc[SC_NLIST] : -1);  // (few slots: the scan is as fast)
    }
    PHASE(15);  // raster
    if (tid == 0) bulk_wait_read0();  // the bulk store must have read the record before the CTA exits
    PHASE(16);  // wait for the bulk store
    if (order_slot >= 0 && order_slot < a.order_stride) a.order_out[8 + order_cls * a.order_stride + order_slot] = arena;
    PHASE_FLUSH;
}

// -------------------------------------------------------------------------------- GAE
// rl.gae / rl.make_returns (rl/__init__.py:94-159; numpy statement rl/test.py:12-30) over a
// time-major [T,N] rollout buffer: one thread per column, reverse scan over t.  The recurrence is
// sequential in t (and must stay so: the float sequence is the reference's), so the kernel is a
// memory-latency problem: the loads of the NEXT batch of U steps are issued before the dependent
// chain of the current batch runs (a ring of NB register batches keeps (NB-1)*U rows in flight per thread).
template <int U>
struct GaeBatch { float r[U], v[U]; uint8_t d[U]; };

template <int U>
__device__ __forceinline__ void gae_load(GaeBatch<U>& b, const float* __restrict__ r, const float* __restrict__ v,
                                         const uint8_t* __restrict__ done, bool want_v, int t1, int N, int n) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int t = t1 - 1 - u;
        const bool ok = t >= 0;
        const size_t i = (size_t)(ok ? t : 0) * N + n;
        b.r[u] = ok ? __ldcs(r + i) : 0.0f;
        b.v[u] = (ok && want_v) ? __ldcs(v + i) : 0.0f;
        b.d[u] = (ok && done) ? __ldcs(done + i) : (uint8_t)0;
    }
}

template <int U>
__device__ __forceinline__ void gae_scan(const GaeBatch<U>& b, float g, float cgl, float* __restrict__ adv,
                                         float* __restrict__ ret, int t1, int T, int N, int n, float& a_next,
                                         float& g_next, float& v_next) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int t = t1 - 1 - u;
        if (t < 0) break;
        const size_t i = (size_t)t * N + n;
        const bool cut = b.d[u] != 0;
        if (adv) {
            float vn = cut ? 0.0f : v_next;
            float delta = (b.r[u] + g * vn) - b.v[u];
            float av = (t == T - 1 || cut) ? delta : delta + cgl * a_next;
            __stcs(adv + i, av);
            a_next = av;
            v_next = b.v[u];
        }
        if (ret) {
            float gn = cut ? 0.0f : g_next;
            float G_t = b.r[u] + g * gn;
            __stcs(ret + i, G_t);
            g_next = G_t;
        }
    }
}

template <int U, int NB>
__global__ void __launch_bounds__(256) agar_gae_kernel(const float* __restrict__ r, const float* __restrict__ v,
                                                      const uint8_t* __restrict__ done,
                                                      const float* __restrict__ end_value, float g, float cgl,
                                                      float* __restrict__ adv, float* __restrict__ ret, int T, int N) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const bool want_v = adv != nullptr;
    // ring of NB register batches: NB-1 batches of loads are in flight while one batch is scanned
    GaeBatch<U> b[NB];
#pragma unroll
    for (int k = 0; k < NB - 1; ++k)
        if (T - k * U > 0) gae_load<U>(b[k], r, v, done, want_v, T - k * U, N, n);
    float a_next = 0.0f;
    float g_next = end_value ? end_value[n] : 0.0f;
    float v_next = want_v ? __ldcs(v + (size_t)T * N + n) : 0.0f;
    for (int t1 = T; t1 > 0; t1 -= NB * U) {
#pragma unroll
        for (int k = 0; k < NB; ++k) {
            const int tcur = t1 - k * U;
            if (tcur <= 0) break;
            const int tpre = tcur - (NB - 1) * U;
            if (tpre > 0) gae_load<U>(b[(k + NB - 1) % NB], r, v, done, want_v, tpre, N, n);
            gae_scan<U>(b[k], g, cgl, adv, ret, tcur, T, N, n, a_next, g_next, v_next);
        }
    }
}

// The same scan fed by TMA: a CTA owns a strip of adjacent columns.  A producer warp streams batches of U rows (4 B of r,
// 4 B of v, 1 B of done per column and row, one bulk copy per row and array) into a ring of STAGES shared-memory stages,
// completion counted on a "full" mbarrier per stage; the W consumer threads scan the stage that has landed, one column
// each, and every consumer warp releases the stage on its "empty" mbarrier, which the producer waits for before it
// refills the stage: no CTA-wide barrier, a warp never waits for another consumer warp.  The strip width is chosen by
// the host so that the strips fill every SM's resident slots evenly (65,536 columns: 586 strips of 112 on 148 SMs x 4;
// with fixed 128-column strips the SMs hold 3 or 4 CTAs and the launch lasts as long as the fuller ones).  The bytes in
// flight do not live in registers, so the ring can be deep (STAGES x U rows per CTA) without costing occupancy.  Needs
// N % 16 == 0 and 16-byte aligned buffers (the host falls back to agar_gae_kernel otherwise).  Block = W + 32 threads.
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
template <int U, int STAGES, int W>
__global__ void __launch_bounds__(W + 32) agar_gae_tma_kernel(const float* __restrict__ r, const float* __restrict__ v,
                                                               const uint8_t* __restrict__ done,
                                                               const float* __restrict__ end_value, float g, float cgl,
                                                               float* __restrict__ adv, float* __restrict__ ret, int T, int N, int strip) {
    __shared__ __align__(16) float sr[STAGES][U][W];
    __shared__ __align__(16) float sv[STAGES][U][W];
    __shared__ __align__(16) uint8_t sd[STAGES][U][W];
    __shared__ __align__(8) uint64_t full[STAGES], empty[STAGES];
    const int tid = threadIdx.x, col0 = blockIdx.x * strip, n = col0 + tid;
    const int width = N - col0 < strip ? N - col0 : strip;  // columns of this strip (a multiple of 16, <= W)
    const bool want_v = adv != nullptr, want_d = done != nullptr;
    const int nb = (T + U - 1) / U;
    static_assert(3 * U <= 32, "one lane of the producer warp per (row, array) of a batch");
    const int lane = tid & 31;
    if (tid == 0) {
        for (int st = 0; st < STAGES; ++st) { mbar_init(&full[st], 1); mbar_init(&empty[st], W / 32); }
    }
    __syncthreads();
    if (tid >= W) {
        // ---- producer warp.  Batch k: rows t = T-1-k*U downwards, at most U of them, into stage k % STAGES: lane 0 posts
        //      the byte count, then lane 3u+w issues the copy of row u of array w (r, v, done) ----
        for (int k = 0; k < nb; ++k) {
            const int st = k % STAGES, t_hi = T - 1 - k * U, rows = t_hi + 1 < U ? t_hi + 1 : U;
            if (k >= STAGES) mbar_wait(&empty[st], (uint32_t)((k / STAGES - 1) & 1));  // the consumers have read batch k - STAGES
            if (lane == 0) mbar_expect_tx(&full[st], (uint32_t)rows * (uint32_t)width * (4u + (want_v ? 4u : 0u) + (want_d ? 1u : 0u)));
            __syncwarp();
            const int u = lane / 3, w = lane - 3 * u;
            if (u < rows && u < U) {
                const size_t off = (size_t)(t_hi - u) * N + col0;
                if (w == 0) bulk_g2s(&sr[st][u][0], r + off, 4u * width, &full[st]);
                else if (w == 1) { if (want_v) bulk_g2s(&sv[st][u][0], v + off, 4u * width, &full[st]); }
                else if (want_d) bulk_g2s(&sd[st][u][0], done + off, (uint32_t)width, &full[st]);
            }
        }
        return;
    }
    // ---- consumers: one column per thread ----
    const bool active = tid < width;
    float a_next = 0.0f;
    float g_next = (end_value && active) ? end_value[n] : 0.0f;
    float v_next = (want_v && active) ? __ldcs(v + (size_t)T * N + n) : 0.0f;
    // one row of the scan; `first` = the row t == T-1 (no successor term)
    auto row = [&](int st, int u, size_t i, bool first) {
        const float rr = sr[st][u][tid];
        const bool cut = want_d && sd[st][u][tid] != 0;
        if (adv) {
            const float vv = sv[st][u][tid];
            const float vn = cut ? 0.0f : v_next;
            const float delta = (rr + g * vn) - vv;
            const float av = (first || cut) ? delta : delta + cgl * a_next;
            __stcs(adv + i, av);
            a_next = av;
            v_next = vv;
        }
        if (ret) {
            const float gn = cut ? 0.0f : g_next;
            const float G_t = rr + g * gn;
            __stcs(ret + i, G_t);
            g_next = G_t;
        }
    };
    for (int k = 0; k < nb; ++k) {
        const int st = k % STAGES, t_hi = T - 1 - k * U, rows = t_hi + 1 < U ? t_hi + 1 : U;
        mbar_wait(&full[st], (uint32_t)((k / STAGES) & 1));
        if (active) {
            size_t i = (size_t)t_hi * N + n;
            if (rows == U) {  // every batch but possibly the last: no per-row bounds
#pragma unroll
                for (int u = 0; u < U; ++u) { row(st, u, i, u == 0 && k == 0); i -= (size_t)N; }
            } else {
                for (int u = 0; u < rows; ++u) { row(st, u, i, u == 0 && k == 0); i -= (size_t)N; }
            }
        }
        __syncwarp();  // the whole warp has read stage st
        if (lane == 0) mbar_arrive(&empty[st]);
    }
}

}  // namespace agar
