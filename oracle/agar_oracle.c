/*
 * agar_oracle.c — CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference`
 * legs may build, load or call this file.  The product (agarai_b200/, libagar_b200.so)
 * never links or calls it and has no CPU fallback.
 *
 * What it restates, and how it is pinned:
 *   - oracle_grid_extract / centroid: features/extractors.py:39-96 and :207-212 of the
 *     reference, evaluated in float32 exactly as numpy 2.3 evaluates them on float32 inputs
 *     (pairwise summation in np.average, weak-scalar casts, int() truncation).  PINNED:
 *     tests/golden/grid_*.npz were produced by importing the reference file itself
 *     (tests/golden/make_golden.py) and tests/test_oracle_golden.py checks this code
 *     against them (integer counts bit-exact, masses to 1e-6 relative).
 *   - oracle_gae / oracle_returns: rl/test.py:12-30 (`normal_gae`, `make_returns`), the
 *     numpy statements the reference's own tests hold rl.gae / rl.make_returns to, and
 *     a2c/losses.py:66-78.  PINNED: known-answer vectors of rl/test.py:59-81 and the
 *     properties of test/returns_test.py:15-46 (tests/test_oracle_golden.py).
 *   - the env step (motion, eat, split, merge, virus, respawn, bots): PARITY UNPINNED.
 *     The reference's simulator is the external, un-vendored, unpinned `gym_agario`
 *     package (imported at ppo/train.py:8, a2c/train.py:8, dqn/train.py:16); it is not
 *     installed, not on disk and no reference test holds any env result.  The rules
 *     implemented here are this build's own spec and this file is their
 *     sequential definition: the CUDA engine must reproduce it bit for bit.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -pthread (oracle/Makefile).  All float
 * arithmetic is single IEEE operations in the order written; no FMA contraction.
 */
#include <math.h>
#include <pthread.h>
#include <stdatomic.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/agar_b200.h"
#include "../include/agar_spec.h"

#define MC AGAR_MAX_CELLS
#define NONE 0x7fffffff

typedef struct OracleEnv {
    AgarConfig cfg;
    AgarLayout L;
    uint64_t seed;
    float* state;    /* [N, L.words] */
    int32_t* events; /* [N, cap, 4]  */
    int32_t* ev_count;
} OracleEnv;

/* ------------------------------------------------------------------ Philox4x32-10 ---- */
/* Salmon et al., "Parallel random numbers: as easy as 1, 2, 3" (SC'11); the Random123
 * constants.  Known-answer vectors are checked in tests/test_oracle_golden.py. */
void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static float u01(uint32_t u) { return (float)(u >> 8) * (1.0f / 16777216.0f); }

static void rand_pos(const OracleEnv* e, int64_t gid, uint32_t slot, uint32_t stream, uint32_t tick,
                     float* x, float* y) {
    uint32_t ctr[4] = {slot, stream, tick, (uint32_t)(uint64_t)gid};
    uint32_t key[2] = {(uint32_t)e->seed, (uint32_t)(e->seed >> 32) ^ (uint32_t)((uint64_t)gid >> 32)};
    uint32_t o[4];
    oracle_philox4x32_10(ctr, key, o);
    *x = u01(o[0]) * e->cfg.arena_size;
    *y = u01(o[1]) * e->cfg.arena_size;
}

/* ------------------------------------------------------------ numpy-order summation --- */
/* numpy's pairwise add-reduce for a contiguous run of n <= 16 float32 values
 * (numpy/core/src/umath/loops_utils.h.src, pairwise_sum: n < 8 sequential from 0; else 8
 * accumulators, combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), remainder sequential).
 * np.average(player[:, (0,1)], axis=0, weights=player[:, -1]) (features/extractors.py:211)
 * reduces all three of sum(w), sum(x*w), sum(y*w) this way (the fancy-indexed copy is
 * F-contiguous).  Verified against numpy 2.3.5 by tests/golden/make_golden.py. */
static float np_sum16(const float* a, int n) {
    if (n < 8) {
        float res = 0.0f;
        for (int i = 0; i < n; ++i) res = res + a[i];
        return res;
    }
    float r[8];
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    int i = 8;
    for (; i < n - (n % 8); i += 8)
        for (int j = 0; j < 8; ++j) r[j] = r[j] + a[i + j];
    float res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res = res + a[i];
    return res;
}

/* features/extractors.py:207-212 `position`: mass-weighted centroid of the alive cells in
 * slot order.  Returns 0 if the player has no cell (reference: None). */
static int centroid16(const float* x, const float* y, const float* m, float* cx, float* cy) {
    float w[MC], xw[MC], yw[MC];
    int n = 0;
    for (int s = 0; s < MC; ++s)
        if (m[s] > 0.0f) { w[n] = m[s]; xw[n] = x[s] * m[s]; yw[n] = y[s] * m[s]; ++n; }
    if (n == 0) return 0;
    float den = np_sum16(w, n);
    *cx = np_sum16(xw, n) / den;
    *cy = np_sum16(yw, n) / den;
    return 1;
}

/* fixed-shape sum of the 16 slots used for rewards: ((m[i]+m[i+8]) ...) butterfly 8,4,2,1 */
static float treesum16(const float* m) {
    float t8[8], t4[4], t2[2];
    for (int i = 0; i < 8; ++i) t8[i] = m[i] + m[i + 8];
    for (int i = 0; i < 4; ++i) t4[i] = t8[i] + t8[i + 4];
    for (int i = 0; i < 2; ++i) t2[i] = t4[i] + t4[i + 2];
    return t2[0] + t2[1];
}

static float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* ------------------------------------------------------------------- grid extractor --- */
/* GridFeatureExtractor.extract (features/extractors.py:39-96) for one agent.
 *   channels in order pellets, own cells, others, viruses, foods as selected by flags.
 *   add_ones (:75-82): gx = int((ex - cx) / box) + G//2, x is the FIRST spatial index;
 *   value = 1 for 2-wide entities, mass for cell rows.
 *   add_out_of_bounds (:84-96): cell (i,j) := -1 when its corner is outside [0, arena).
 *   The `others` channel gets its sentinel pass only if there is at least one other
 *   player (:53-56 calls add_entities_to_grid once per other player).
 * out: float[C*G*G].  Returns 0 (and an all-zero observation) if the agent has no cell. */
static void add_point(float* ch, int G, float box, float cx, float cy, float ex, float ey, float val) {
    int gx = (int)((ex - cx) / box) + G / 2;
    int gy = (int)((ey - cy) / box) + G / 2;
    if (0 <= gx && gx < G && 0 <= gy && gy < G) ch[gx * G + gy] += val;
}

static void add_oob(float* ch, int G, double box_d, float cx, float cy, float arena) {
    for (int i = 0; i < G; ++i)
        for (int j = 0; j < G; ++j) {
            float xl = (float)((double)(i - G / 2) * box_d) + cx;
            float yl = (float)((double)(j - G / 2) * box_d) + cy;
            if (!(0.0f <= xl && xl < arena && 0.0f <= yl && yl < arena)) ch[i * G + j] = -1.0f;
        }
}

int oracle_grid_extract(int G, float view, float arena, int flags, int agent, int K,
                        const float* pel_x, const float* pel_y, int P,
                        const float* vir_x, const float* vir_y, int V,
                        const float* food_x, const float* food_y, int F,
                        const float* cell_x, const float* cell_y, const float* cell_m, /* [K*16] */
                        float* out) {
    int C = 0;
    for (int b = 0; b < 5; ++b) C += (flags >> b) & 1;
    memset(out, 0, sizeof(float) * (size_t)C * G * G);
    float cx, cy;
    const float* ax = cell_x + agent * MC; const float* ay = cell_y + agent * MC;
    const float* am = cell_m + agent * MC;
    if (!centroid16(ax, ay, am, &cx, &cy)) return 0;
    double box_d = (double)view / (double)G;
    float box = (float)box_d;
    int c = 0;
    if (flags & AGAR_OBS_PELLETS) {
        float* ch = out + (size_t)c * G * G;
        for (int p = 0; p < P; ++p)
            if (pel_x[p] >= 0.0f) add_point(ch, G, box, cx, cy, pel_x[p], pel_y[p], 1.0f);
        add_oob(ch, G, box_d, cx, cy, arena);
        ++c;
    }
    if (flags & AGAR_OBS_CELLS) {
        float* ch = out + (size_t)c * G * G;
        for (int s = 0; s < MC; ++s)
            if (am[s] > 0.0f) add_point(ch, G, box, cx, cy, ax[s], ay[s], am[s]);
        add_oob(ch, G, box_d, cx, cy, arena);
        ++c;
    }
    if (flags & AGAR_OBS_OTHERS) {
        float* ch = out + (size_t)c * G * G;
        for (int k = 0; k < K; ++k) {
            if (k == agent) continue;
            for (int s = 0; s < MC; ++s) {
                int g = k * MC + s;
                if (cell_m[g] > 0.0f) add_point(ch, G, box, cx, cy, cell_x[g], cell_y[g], cell_m[g]);
            }
        }
        if (K > 1) add_oob(ch, G, box_d, cx, cy, arena);
        ++c;
    }
    if (flags & AGAR_OBS_VIRUSES) {
        float* ch = out + (size_t)c * G * G;
        for (int v = 0; v < V; ++v) add_point(ch, G, box, cx, cy, vir_x[v], vir_y[v], 1.0f);
        add_oob(ch, G, box_d, cx, cy, arena);
        ++c;
    }
    if (flags & AGAR_OBS_FOODS) {
        float* ch = out + (size_t)c * G * G;
        for (int f = 0; f < F; ++f)
            if (food_x[f] >= 0.0f) add_point(ch, G, box, cx, cy, food_x[f], food_y[f], 1.0f);
        add_oob(ch, G, box_d, cx, cy, arena);
        ++c;
    }
    return 1;
}

/* centroid exposed for the golden check of `position` (features/extractors.py:207-212) */
int oracle_centroid(const float* x, const float* y, const float* m, float* out_xy) {
    return centroid16(x, y, m, &out_xy[0], &out_xy[1]);
}

/* --------------------------------------------------------------------- GAE / returns --- */
/* rl/test.py:12-21 normal_gae and :24-30 make_returns, one column of a time-major [T,N]
 * buffer per agent.  c = float32(gamma*lam) with the product taken in double, as numpy
 * does for `(gamma * lam) * adv[t + 1]` on float32 data.  done (optional, not in the
 * reference): done[t] != 0 cuts the recursion after step t. */
void oracle_gae(const float* r, const float* v, const uint8_t* done, const float* end_value,
                double gamma, double lam, float* adv, float* ret, int T, int N) {
    float g = (float)gamma, c = (float)(gamma * lam);
    for (int n = 0; n < N; ++n) {
        float a_next = 0.0f;
        float g_next = end_value ? end_value[n] : 0.0f;
        for (int t = T - 1; t >= 0; --t) {
            size_t i = (size_t)t * N + n;
            int cut = done && done[i];
            if (adv) {
                float vn = cut ? 0.0f : v[i + N];
                float delta = (r[i] + g * vn) - v[i];
                float a = (t == T - 1 || cut) ? delta : delta + c * a_next;
                adv[i] = a;
                a_next = a;
            }
            if (ret) {
                float gn = cut ? 0.0f : g_next;
                float G_t = r[i] + g * gn;
                ret[i] = G_t;
                g_next = G_t;
            }
        }
    }
}

/* ----------------------------------------------------------------------------- env --- */
static void ev_push(OracleEnv* e, int n, int tick, int type, int a, int b) {
    int cap = e->cfg.event_capacity;
    if (cap <= 0) return;
    int i = e->ev_count[n]++;
    if (i < cap) {
        int32_t* p = e->events + ((size_t)n * cap + i) * 4;
        p[0] = tick; p[1] = type; p[2] = a; p[3] = b;
    }
}

/* indices of the alive cells in ascending gid order (so "first hit" is the lowest gid) */
static int alive_list(const float* cm, int KC, int* out) {
    int n = 0;
    for (int g = 0; g < KC; ++g)
        if (cm[g] > 0.0f) out[n++] = g;
    return n;
}

static void clear_cell(float* S, const AgarLayout* L, int g) {
    S[L->cell_x + g] = 0; S[L->cell_y + g] = 0; S[L->cell_ix + g] = 0;
    S[L->cell_iy + g] = 0; S[L->cell_m + g] = 0; S[L->cell_t + g] = 0;
}

static void spawn_player(OracleEnv* e, float* S, int64_t gid, int k, uint32_t tick) {
    const AgarLayout* L = &e->L;
    for (int s = 0; s < MC; ++s) clear_cell(S, L, k * MC + s);
    float x, y;
    rand_pos(e, gid, (uint32_t)k, AGAR_STREAM_PLAYER_SPAWN, tick, &x, &y);
    S[L->cell_x + k * MC] = x; S[L->cell_y + k * MC] = y;
    S[L->cell_m + k * MC] = e->cfg.start_mass;
}

static void reset_arena(OracleEnv* e, int n) {
    const AgarConfig* c = &e->cfg; const AgarLayout* L = &e->L;
    float* S = e->state + (size_t)n * L->words;
    int32_t* Si = (int32_t*)S;
    int64_t gid = c->arena_id_base + n;
    uint32_t tick = (uint32_t)Si[L->tick];
    memset(S, 0, sizeof(float) * (size_t)L->words);
    Si[L->tick] = (int32_t)tick;
    for (int p = 0; p < c->num_pellets; ++p)
        rand_pos(e, gid, (uint32_t)p, AGAR_STREAM_PELLET_INIT, tick, &S[L->pellet_x + p], &S[L->pellet_y + p]);
    for (int p = c->num_pellets; p < L->P_pad; ++p) { S[L->pellet_x + p] = -1.0f; S[L->pellet_y + p] = -1.0f; }
    for (int v = 0; v < c->num_viruses; ++v)
        rand_pos(e, gid, (uint32_t)v, AGAR_STREAM_VIRUS_INIT, tick, &S[L->virus_x + v], &S[L->virus_y + v]);
    for (int v = c->num_viruses; v < L->V_pad; ++v) { S[L->virus_x + v] = -1.0f; S[L->virus_y + v] = -1.0f; }
    for (int f = 0; f < L->F_pad; ++f) S[L->food_x + f] = -1.0f;
    for (int k = 0; k < L->num_players; ++k) spawn_player(e, S, gid, k, tick);
}

/* one arena, one env step (ticks_per_step ticks).  DESIGN.md §3 is the prose of this. */
static void step_arena(OracleEnv* e, int n, const float* target_xy, const int32_t* act,
                       float* reward, uint8_t* done) {
    const AgarConfig* c = &e->cfg; const AgarLayout* L = &e->L;
    const int A = c->num_agents, K = L->num_players, KC = L->cells_total;
    const int P = c->num_pellets, V = c->num_viruses, F = c->max_foods;
    float* S = e->state + (size_t)n * L->words;
    int32_t* Si = (int32_t*)S;
    int64_t gid = c->arena_id_base + n;
    float *cx = S + L->cell_x, *cy = S + L->cell_y, *cix = S + L->cell_ix, *ciy = S + L->cell_iy;
    float *cm = S + L->cell_m, *ct = S + L->cell_t;
    float *px = S + L->pellet_x, *py = S + L->pellet_y, *vx = S + L->virus_x, *vy = S + L->virus_y;
    float *fx = S + L->food_x, *fy = S + L->food_y, *fvx = S + L->food_vx, *fvy = S + L->food_vy;
    const float arena = c->arena_size, r2pm = c->r2_per_mass;
    static const float DX[16] = AGAR_DIR16_X, DY[16] = AGAR_DIR16_Y;
    if (c->event_capacity > 0) e->ev_count[n] = 0;

    float mass_before[64];
    float Tx[64], Ty[64], dirx[64], diry[64];
    int pact[64];
    uint32_t tick0 = (uint32_t)Si[L->tick];

    for (int a = 0; a < A; ++a) mass_before[a] = treesum16(cm + a * MC);

    /* ---- step start: bot respawn, targets ---- */
    for (int k = 0; k < K; ++k) {
        int alive = 0;
        for (int s = 0; s < MC; ++s) alive |= cm[k * MC + s] > 0.0f;
        pact[k] = 0; Tx[k] = Ty[k] = 0.0f; dirx[k] = 1.0f; diry[k] = 0.0f;
        if (k >= A && !alive) {
            spawn_player(e, S, gid, k, tick0);
            ev_push(e, n, 0, AGAR_EV_RESPAWN, k, 0);
            alive = 1;
        }
        if (!alive) continue;
        float Cx, Cy;
        centroid16(cx + k * MC, cy + k * MC, cm + k * MC, &Cx, &Cy);
        if (k < A) {
            float ax = target_xy[((size_t)n * A + k) * 2 + 0], ay = target_xy[((size_t)n * A + k) * 2 + 1];
            float n2 = ax * ax + ay * ay;
            if (!(n2 < INFINITY)) { ax = 0.0f; ay = 0.0f; }
            else if (n2 > 1.0f) { float inv = 1.0f / sqrtf(n2); ax = ax * inv; ay = ay * inv; }
            Tx[k] = Cx + ax * c->target_reach;
            Ty[k] = Cy + ay * c->target_reach;
            float m2 = ax * ax + ay * ay;
            if (m2 > 0.0f) { float inv = 1.0f / sqrtf(m2); dirx[k] = ax * inv; diry[k] = ay * inv; }
            int a_ = act[(size_t)n * A + k];
            pact[k] = (a_ == AGAR_ACT_SPLIT || a_ == AGAR_ACT_FEED) ? a_ : 0;
        } else {
            /* bot: nearest pellet inside the square sense window, else roam way-point */
            float R = c->bot_sense_radius, best = INFINITY; int bi = -1;
            for (int p = 0; p < P; ++p) {
                if (!(px[p] >= 0.0f)) continue;
                float dx = px[p] - Cx, dy = py[p] - Cy;
                if (fabsf(dx) <= R && fabsf(dy) <= R) {
                    float d2 = dx * dx + dy * dy;
                    if (d2 < best) { best = d2; bi = p; }
                }
            }
            if (bi >= 0) { Tx[k] = px[bi]; Ty[k] = py[bi]; }
            else {
                uint32_t epoch = tick0 & ~(uint32_t)(c->bot_roam_period - 1);
                rand_pos(e, gid, (uint32_t)k, AGAR_STREAM_BOT_ROAM, epoch, &Tx[k], &Ty[k]);
            }
        }
    }
    /* ---- split (act 1): i-th eligible cell -> i-th free slot, on the pre-split state ---- */
    for (int k = 0; k < A; ++k) {
        if (pact[k] != AGAR_ACT_SPLIT) continue;
        int elig[MC], nfree = 0, ne = 0, freeslot[MC];
        for (int s = 0; s < MC; ++s) {
            float m = cm[k * MC + s];
            if (m == 0.0f) freeslot[nfree++] = s;
            else if (m >= c->split_min_mass) elig[ne++] = s;
        }
        int np = ne < nfree ? ne : nfree;
        for (int i = 0; i < np; ++i) {
            int src = k * MC + elig[i], dst = k * MC + freeslot[i];
            float half = cm[src] * 0.5f;
            cm[src] = half; ct[src] = c->merge_delay;
            cx[dst] = cx[src]; cy[dst] = cy[src];
            cix[dst] = dirx[k] * c->split_speed; ciy[dst] = diry[k] * c->split_speed;
            cm[dst] = half; ct[dst] = c->merge_delay;
            ev_push(e, n, 0, AGAR_EV_SPLIT, src, dst);
        }
    }
    /* ---- feed (act 2): cells in gid order take free food slots in ascending order ---- */
    {
        int f = 0;
        for (int k = 0; k < A; ++k) {
            if (pact[k] != AGAR_ACT_FEED) continue;
            for (int s = 0; s < MC; ++s) {
                int g = k * MC + s;
                if (!(cm[g] >= c->feed_min_mass)) continue;
                while (f < F && fx[f] >= 0.0f) ++f;
                if (f >= F) break;
                float r = sqrtf(cm[g] * r2pm);
                cm[g] = cm[g] - c->food_mass;
                float off = r + c->food_spawn_gap;
                fx[f] = clampf(cx[g] + dirx[k] * off, 0.0f, arena);
                fy[f] = clampf(cy[g] + diry[k] * off, 0.0f, arena);
                fvx[f] = dirx[k] * c->feed_speed; fvy[f] = diry[k] * c->feed_speed;
                ev_push(e, n, 0, AGAR_EV_FEED, g, f);
                ++f;
            }
        }
    }

    int* win = (int*)malloc(sizeof(int) * (size_t)(2 * KC + V + 8));
    int* alive = win + KC + V + 8; int na;
    float* snap_m = (float*)malloc(sizeof(float) * (size_t)KC * 2);
    float* gain = snap_m + KC;

    for (int tk = 0; tk < c->ticks_per_step; ++tk) {
        uint32_t tick = tick0 + (uint32_t)tk;
        /* ---- motion ---- */
        for (int g = 0; g < KC; ++g) {
            float m = cm[g];
            if (!(m > 0.0f)) continue;
            if (c->mass_decay > 0.0f) { /* mass decay: the cell keeps (1 - mass_decay) of its mass, never below start_mass */
                float md = m * (1.0f - c->mass_decay);
                if (md >= c->start_mass) { m = md; cm[g] = m; }
            }
            int k = g / MC;
            float r = sqrtf(m * r2pm);
            float spd = c->speed_k / sqrtf(r);
            float dx = Tx[k] - cx[g], dy = Ty[k] - cy[g];
            float d = sqrtf(dx * dx + dy * dy);
            float den = d > c->target_reach ? d : c->target_reach;
            float s = spd / den;
            float nx = cx[g] + (dx * s + cix[g]);
            float ny = cy[g] + (dy * s + ciy[g]);
            float ix = cix[g] * c->impulse_decay, iy = ciy[g] * c->impulse_decay;
            if (fabsf(ix) < 1e-3f && fabsf(iy) < 1e-3f) { ix = 0.0f; iy = 0.0f; }
            cx[g] = clampf(nx, 0.0f, arena); cy[g] = clampf(ny, 0.0f, arena);
            cix[g] = ix; ciy[g] = iy;
            if (ct[g] > 0.0f) ct[g] = ct[g] - 1.0f;
        }
        /* ---- pellets: eaten by the lowest-gid cell whose disc holds the pellet centre ---- */
        for (int g = 0; g < KC; ++g) { snap_m[g] = cm[g]; win[g] = 0; }
        na = alive_list(cm, KC, alive);
        for (int p = 0; p < P; ++p) {
            if (!(px[p] >= 0.0f)) continue;
            for (int q = 0; q < na; ++q) {
                int g = alive[q];
                float dx = px[p] - cx[g], dy = py[p] - cy[g];
                if (dx * dx + dy * dy < snap_m[g] * r2pm) {
                    win[g]++;
                    ev_push(e, n, tk, AGAR_EV_PELLET_EATEN, g, p);
                    if (c->pellet_regen) rand_pos(e, gid, (uint32_t)p, AGAR_STREAM_PELLET_RESPAWN, tick, &px[p], &py[p]);
                    else { px[p] = -1.0f; py[p] = -1.0f; }
                    break;
                }
            }
        }
        for (int g = 0; g < KC; ++g)
            if (win[g]) cm[g] = cm[g] + (float)win[g] * c->pellet_mass;
        /* ---- foods: move, then eaten like pellets ---- */
        for (int g = 0; g < KC; ++g) { snap_m[g] = cm[g]; win[g] = 0; }
        na = alive_list(cm, KC, alive);
        for (int f = 0; f < F; ++f) {
            if (!(fx[f] >= 0.0f)) continue;
            float nx = fx[f] + fvx[f], ny = fy[f] + fvy[f];
            float ux = fvx[f] * c->food_decay, uy = fvy[f] * c->food_decay;
            if (fabsf(ux) < 1e-3f && fabsf(uy) < 1e-3f) { ux = 0.0f; uy = 0.0f; }
            fx[f] = clampf(nx, 0.0f, arena); fy[f] = clampf(ny, 0.0f, arena);
            fvx[f] = ux; fvy[f] = uy;
            for (int q = 0; q < na; ++q) {
                int g = alive[q];
                float dx = fx[f] - cx[g], dy = fy[f] - cy[g];
                if (dx * dx + dy * dy < snap_m[g] * r2pm) {
                    win[g]++;
                    ev_push(e, n, tk, AGAR_EV_FOOD_EATEN, g, f);
                    fx[f] = -1.0f; fy[f] = 0.0f; fvx[f] = 0.0f; fvy[f] = 0.0f;
                    break;
                }
            }
        }
        for (int g = 0; g < KC; ++g)
            if (win[g]) cm[g] = cm[g] + (float)win[g] * c->food_mass;
        /* ---- viruses ---- */
        if (V > 0) {
            int* vwin = win + KC; /* [V] winner cell of each virus */
            for (int g = 0; g < KC; ++g) win[g] = NONE; /* first virus won by each cell */
            na = 0;
            for (int g = 0; g < KC; ++g)
                if (cm[g] >= c->virus_pop_mass) alive[na++] = g; /* heavy enough to pop */
            for (int v = 0; v < V; ++v) {
                vwin[v] = NONE;
                for (int q = 0; q < na; ++q) {
                    int g = alive[q];
                    float dx = vx[v] - cx[g], dy = vy[v] - cy[g];
                    if (dx * dx + dy * dy < cm[g] * r2pm) { vwin[v] = g; break; }
                }
                if (vwin[v] != NONE && win[vwin[v]] == NONE) win[vwin[v]] = v;
            }
            for (int g = 0; g < KC; ++g) {
                int v = win[g];
                if (v == NONE) continue;
                int k = g / MC;
                float m1 = cm[g] + c->virus_mass;
                int freeslot[MC], nfree = 0;
                for (int s = 0; s < MC; ++s)
                    if (cm[k * MC + s] == 0.0f) freeslot[nfree++] = s;
                int np = nfree < c->pop_pieces ? nfree : c->pop_pieces;
                ev_push(e, n, tk, AGAR_EV_VIRUS_POP, g, v);
                if (np > 0) {
                    float keep = m1 * 0.5f;
                    float piece = (m1 - keep) / (float)np;
                    cm[g] = keep; ct[g] = c->merge_delay;
                    for (int j = 0; j < np; ++j) {
                        int d = k * MC + freeslot[j];
                        cx[d] = cx[g]; cy[d] = cy[g];
                        cix[d] = DX[j] * c->pop_speed; ciy[d] = DY[j] * c->pop_speed;
                        cm[d] = piece; ct[d] = c->merge_delay;
                        ev_push(e, n, tk, AGAR_EV_POP_PIECE, g, d);
                    }
                } else {
                    cm[g] = m1;
                }
                rand_pos(e, gid, (uint32_t)v, AGAR_STREAM_VIRUS_RESPAWN, tick, &vx[v], &vy[v]);
            }
        }
        /* ---- cell eats cell (different players), all on the pre-phase snapshot ---- */
        if (K > 1) {
            for (int g = 0; g < KC; ++g) { snap_m[g] = cm[g]; gain[g] = 0.0f; win[g] = NONE; }
            na = alive_list(cm, KC, alive);
            for (int qj = 0; qj < na; ++qj) {
                int j = alive[qj];
                for (int qi = 0; qi < na; ++qi) {
                    int i = alive[qi];
                    if (i / MC == j / MC) continue;
                    if (!(snap_m[i] >= snap_m[j] * c->eat_ratio)) continue;
                    float dx = cx[j] - cx[i], dy = cy[j] - cy[i];
                    if (dx * dx + dy * dy < snap_m[i] * r2pm) { win[j] = i; break; }
                }
            }
            for (int j = 0; j < KC; ++j)
                if (win[j] != NONE) {
                    gain[win[j]] = gain[win[j]] + snap_m[j];
                    ev_push(e, n, tk, AGAR_EV_CELL_EATEN, win[j], j);
                }
            for (int g = 0; g < KC; ++g) {
                if (win[g] != NONE) clear_cell(S, L, g);       /* eaten: forfeits its own gains */
                else if (gain[g] > 0.0f) cm[g] = cm[g] + gain[g];
            }
        }
        /* ---- merge own cells whose timers ran out ---- */
        for (int k = 0; k < K; ++k) {
            int root[MC]; float sm[MC];
            int any = 0;
            for (int s = 0; s < MC; ++s) { sm[s] = cm[k * MC + s]; root[s] = s; }
            for (int b = 1; b < MC; ++b) {
                int gb = k * MC + b;
                if (!(sm[b] > 0.0f) || ct[gb] != 0.0f) continue;
                for (int a = 0; a < b; ++a) {
                    int ga = k * MC + a;
                    if (!(sm[a] > 0.0f) || ct[ga] != 0.0f) continue;
                    float dx = cx[gb] - cx[ga], dy = cy[gb] - cy[ga];
                    float ra = sm[a] * r2pm, rb = sm[b] * r2pm;
                    float rr = ra > rb ? ra : rb;
                    if (dx * dx + dy * dy < rr) { root[b] = root[a]; any = 1; break; }
                }
            }
            if (!any) continue;
            for (int s = 1; s < MC; ++s) {
                if (root[s] == s) continue;
                int gr = k * MC + root[s];
                cm[gr] = cm[gr] + sm[s];
                ev_push(e, n, tk, AGAR_EV_MERGE, gr, k * MC + s);
                clear_cell(S, L, k * MC + s);
            }
        }
    }
    free(win); free(snap_m);
    Si[L->tick] = (int32_t)(tick0 + (uint32_t)c->ticks_per_step);
    Si[L->episode_steps] += 1;

    /* ---- rewards / dones ---- */
    int all_done = 1;
    for (int a = 0; a < A; ++a) {
        int was_done = Si[L->agent_done + a];
        float after = treesum16(cm + a * MC);
        int alive = 0;
        for (int s = 0; s < MC; ++s) alive |= cm[a * MC + s] > 0.0f;
        float rw = was_done ? 0.0f : after - mass_before[a];
        if (!was_done && !alive) {
            Si[L->agent_done + a] = 1;
            ev_push(e, n, c->ticks_per_step, AGAR_EV_DEATH, a, 0);
        }
        reward[(size_t)n * A + a] = rw;
        done[(size_t)n * A + a] = (uint8_t)Si[L->agent_done + a];
        all_done &= Si[L->agent_done + a];
    }
    if (c->auto_reset && all_done) reset_arena(e, n);
}

static void observe_arena(const OracleEnv* e, const float* S, float* obs /* [A,C,G,G] */) {
    const AgarConfig* c = &e->cfg; const AgarLayout* L = &e->L;
    int C = agar_spec_obs_channels(c), G = c->grid_size;
    for (int a = 0; a < c->num_agents; ++a)
        oracle_grid_extract(G, c->view_size, c->arena_size, c->obs_flags, a, L->num_players,
                            S + L->pellet_x, S + L->pellet_y, c->num_pellets,
                            S + L->virus_x, S + L->virus_y, c->num_viruses,
                            S + L->food_x, S + L->food_y, c->max_foods,
                            S + L->cell_x, S + L->cell_y, S + L->cell_m,
                            obs + (size_t)a * C * G * G);
}


/* ------------------------------------------------------- RAM-style feature extractor --- */
/* FeatureExtractor.extract (features/extractors.py:99-149) and its helpers (:171-212) for one
 * agent, on float32 inputs (numpy then evaluates every step below in float32):
 *   loc = position(agent)                                   (:207-212, centroid16 above)
 *   get_entity_features(loc, E, n)  (:171-180): rows of E ordered by
 *       np.linalg.norm(E[:, :2] - loc, axis=1) = sqrtf(dx*dx + dy*dy), first n, (x, y) -= loc,
 *       zero rows pad.  numpy's default argsort leaves the order of equal keys unspecified; this
 *       restatement (and the CUDA kernel) keep index order (the stable sort).
 *   largest_cells(player, n)        (:182-184): argsort ASCENDING by mass, first n rows - the
 *       n smallest cells (reference quirk, reproduced); stable.
 *   closest_players(loc, others, n) (:186-193): by sqrtf(dx*dx + dy*dy) between centroids; stable.
 *       `others` = the other players that have at least one cell, in player order (own spec).
 *   missing players are filled with -1000 rows (:139-141, filler_value :110).
 * Cell rows are (x, y, impulse_x, impulse_y, mass): the reference never reads columns 2-3.
 * out: float[2*(np+nv+nf) + 5*(1+no)*nc].  Returns 0 and zeros if the agent has no cell (the
 * reference returns None). */
typedef struct { float d; int i; } DistIdx;

static void stable_sort_dist(DistIdx* a, int n) {  /* insertion sort: stable, n is small or moderate */
    for (int i = 1; i < n; ++i) {
        DistIdx v = a[i];
        int j = i - 1;
        while (j >= 0 && a[j].d > v.d) { a[j + 1] = a[j]; --j; }
        a[j + 1] = v;
    }
}

static void ram_points(const float* ex, const float* ey, int count, int marker, float lx, float ly, int n, float* out) {
    DistIdx* a = (DistIdx*)malloc(sizeof(DistIdx) * (size_t)(count > 0 ? count : 1));
    int m = 0;
    for (int i = 0; i < count; ++i) {
        if (marker && !(ex[i] >= 0.0f)) continue;
        float dx = ex[i] - lx, dy = ey[i] - ly;
        float s0 = dx * dx, s1 = dy * dy;
        a[m].d = sqrtf(s0 + s1); a[m].i = i; ++m;
    }
    stable_sort_dist(a, m);
    for (int r = 0; r < n; ++r) {
        if (r < m) { out[2 * r] = ex[a[r].i] - lx; out[2 * r + 1] = ey[a[r].i] - ly; }
        else { out[2 * r] = 0.0f; out[2 * r + 1] = 0.0f; }
    }
    free(a);
}

static void ram_cells(const float* cx, const float* cy, const float* cix, const float* ciy, const float* cm,
                      float lx, float ly, int relative, int n, float* out) {
    DistIdx bym[MC]; int na = 0;
    for (int s = 0; s < MC; ++s)
        if (cm[s] > 0.0f) { bym[na].d = cm[s]; bym[na].i = s; ++na; }
    stable_sort_dist(bym, na);                    /* largest_cells: ascending by mass */
    int nsel = na < n ? na : n;
    DistIdx byd[MC];
    for (int r = 0; r < nsel; ++r) {
        int s = bym[r].i;
        float dx = cx[s] - lx, dy = cy[s] - ly;
        float s0 = dx * dx, s1 = dy * dy;
        byd[r].d = sqrtf(s0 + s1); byd[r].i = s;
    }
    stable_sort_dist(byd, nsel);                  /* sort_by_proximity on the selected rows */
    for (int r = 0; r < n; ++r) {
        float* o = out + 5 * r;
        if (r < nsel) {
            int s = byd[r].i;
            o[0] = relative ? cx[s] - lx : cx[s]; o[1] = relative ? cy[s] - ly : cy[s];
            o[2] = cix[s]; o[3] = ciy[s]; o[4] = cm[s];
        } else { o[0] = o[1] = o[2] = o[3] = o[4] = 0.0f; }
    }
}

int oracle_ram_feature_size(int np_, int nv, int nf, int no, int nc) { return 2 * (np_ + nv + nf) + 5 * (1 + no) * nc; }

int oracle_ram_extract(int n_pellet, int n_virus, int n_food, int n_other, int n_cell, int agent, int K,
                       const float* pel_x, const float* pel_y, int P,
                       const float* vir_x, const float* vir_y, int V,
                       const float* food_x, const float* food_y, int F,
                       const float* cell_x, const float* cell_y, const float* cell_ix, const float* cell_iy,
                       const float* cell_m, /* [K*16] */
                       float* out) {
    int size = oracle_ram_feature_size(n_pellet, n_virus, n_food, n_other, n_cell);
    memset(out, 0, sizeof(float) * (size_t)size);
    float lx, ly;
    if (!centroid16(cell_x + agent * MC, cell_y + agent * MC, cell_m + agent * MC, &lx, &ly)) return 0;
    float* o = out;
    ram_points(pel_x, pel_y, P, 1, lx, ly, n_pellet, o); o += 2 * n_pellet;
    ram_points(vir_x, vir_y, V, 0, lx, ly, n_virus, o); o += 2 * n_virus;
    ram_points(food_x, food_y, F, 1, lx, ly, n_food, o); o += 2 * n_food;
    int g = agent * MC;
    ram_cells(cell_x + g, cell_y + g, cell_ix + g, cell_iy + g, cell_m + g, lx, ly, 0, n_cell, o); o += 5 * n_cell;
    DistIdx pl[64]; int npl = 0;
    for (int k = 0; k < K; ++k) {
        if (k == agent) continue;
        float qx, qy;
        if (!centroid16(cell_x + k * MC, cell_y + k * MC, cell_m + k * MC, &qx, &qy)) continue;
        float dx = lx - qx, dy = ly - qy;
        float s0 = dx * dx, s1 = dy * dy;
        pl[npl].d = sqrtf(s0 + s1); pl[npl].i = k; ++npl;
    }
    stable_sort_dist(pl, npl);
    for (int j = 0; j < n_other; ++j) {
        if (j < npl) {
            int g2 = pl[j].i * MC;
            ram_cells(cell_x + g2, cell_y + g2, cell_ix + g2, cell_iy + g2, cell_m + g2, lx, ly, 1, n_cell, o);
        } else {
            for (int r = 0; r < 5 * n_cell; ++r) o[r] = -1000.0f;
        }
        o += 5 * n_cell;
    }
    return 1;
}

/* FeatureExtractor for every agent of every arena of the env's current state: float[N, A, size] */
void oracle_observe_ram(const OracleEnv* e, int n_pellet, int n_virus, int n_food, int n_other, int n_cell, float* out) {
    const AgarConfig* c = &e->cfg; const AgarLayout* L = &e->L;
    int size = oracle_ram_feature_size(n_pellet, n_virus, n_food, n_other, n_cell);
    for (int n = 0; n < c->num_arenas; ++n) {
        const float* S = e->state + (size_t)n * L->words;
        for (int a = 0; a < c->num_agents; ++a)
            oracle_ram_extract(n_pellet, n_virus, n_food, n_other, n_cell, a, L->num_players,
                               S + L->pellet_x, S + L->pellet_y, c->num_pellets, S + L->virus_x, S + L->virus_y, c->num_viruses,
                               S + L->food_x, S + L->food_y, c->max_foods, S + L->cell_x, S + L->cell_y, S + L->cell_ix,
                               S + L->cell_iy, S + L->cell_m, out + ((size_t)n * c->num_agents + a) * size);
    }
}

/* ------------------------------------------------------------- screen -> grayscale --- */
/* ScreenFeatureExtractor.extract (features/extractors.py:152-164): observation.mean(axis=-1) of
 * uint8 frames: numpy adds the three bytes in float64 and divides by the count. */
void oracle_screen_gray(const uint8_t* rgb, long long n_pixels, double* out) {
    for (long long p = 0; p < n_pixels; ++p)
        out[p] = ((double)rgb[3 * p] + (double)rgb[3 * p + 1] + (double)rgb[3 * p + 2]) / 3.0;
}

/* ---------------------------------------------------------------------- public API --- */
OracleEnv* oracle_create(const AgarConfig* cfg) {
    if (agar_spec_config_check(cfg)) return NULL;
    OracleEnv* e = (OracleEnv*)calloc(1, sizeof(OracleEnv));
    e->cfg = *cfg;
    agar_spec_layout(cfg, &e->L);
    e->state = (float*)calloc((size_t)cfg->num_arenas * e->L.words, sizeof(float));
    if (cfg->event_capacity > 0) {
        e->events = (int32_t*)calloc((size_t)cfg->num_arenas * cfg->event_capacity * 4, sizeof(int32_t));
        e->ev_count = (int32_t*)calloc((size_t)cfg->num_arenas, sizeof(int32_t));
    }
    return e;
}

void oracle_destroy(OracleEnv* e) {
    if (!e) return;
    free(e->state); free(e->events); free(e->ev_count); free(e);
}

int oracle_words(const OracleEnv* e) { return e->L.words; }

void oracle_seed(OracleEnv* e, uint64_t seed) {
    e->seed = seed;
    for (int n = 0; n < e->cfg.num_arenas; ++n) ((int32_t*)(e->state + (size_t)n * e->L.words))[e->L.tick] = 0;
}

/* arenas are independent: a tiny pthread parallel-for hands out blocks of 8 arenas */
typedef struct ParJob {
    OracleEnv* e; int mode; /* 0 observe, 1 reset, 2 step */
    const uint8_t* mask; const float* target_xy; const int32_t* act;
    float* obs; float* reward; uint8_t* done; size_t per;
    atomic_int next;
} ParJob;

static void* par_worker(void* arg) {
    ParJob* j = (ParJob*)arg; OracleEnv* e = j->e;
    int N = e->cfg.num_arenas;
    for (;;) {
        int b = atomic_fetch_add(&j->next, 8);
        if (b >= N) break;
        int hi = b + 8 < N ? b + 8 : N;
        for (int n = b; n < hi; ++n) {
            float* S = e->state + (size_t)n * e->L.words;
            if (j->mode == 1) { if (!j->mask || j->mask[n]) reset_arena(e, n); }
            if (j->mode == 2) step_arena(e, n, j->target_xy, j->act, j->reward, j->done);
            if (j->mode != 1 && j->obs) observe_arena(e, S, j->obs + (size_t)n * j->per);
        }
    }
    return NULL;
}

static void par_run(ParJob* j, int nthreads) {
    atomic_init(&j->next, 0);
    if (nthreads <= 1) { par_worker(j); return; }
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256];
    int started = 0;
    for (int i = 0; i < nthreads - 1; ++i)
        if (pthread_create(&th[started], NULL, par_worker, j) == 0) ++started;
    par_worker(j);
    for (int i = 0; i < started; ++i) pthread_join(th[i], NULL);
}

static size_t obs_per_arena(const OracleEnv* e) {
    return (size_t)e->cfg.num_agents * agar_spec_obs_channels(&e->cfg) * e->cfg.grid_size * e->cfg.grid_size;
}

void oracle_observe(OracleEnv* e, float* obs, int nthreads) {
    ParJob j; memset(&j, 0, sizeof(j));
    j.e = e; j.mode = 0; j.obs = obs; j.per = obs_per_arena(e);
    par_run(&j, nthreads);
}

void oracle_reset(OracleEnv* e, const uint8_t* mask, float* obs, int nthreads) {
    ParJob j; memset(&j, 0, sizeof(j));
    j.e = e; j.mode = 1; j.mask = mask;
    par_run(&j, nthreads);
    if (obs) oracle_observe(e, obs, nthreads);
}

/* env.step for every arena, spread over nthreads host threads (this is the CPU baseline
 * timed in bench.py). */
void oracle_step(OracleEnv* e, const float* target_xy, const int32_t* act, float* obs,
                 float* reward, uint8_t* done, int nthreads) {
    ParJob j; memset(&j, 0, sizeof(j));
    j.e = e; j.mode = 2; j.target_xy = target_xy; j.act = act; j.obs = obs;
    j.reward = reward; j.done = done; j.per = obs_per_arena(e);
    par_run(&j, nthreads);
}

void oracle_get_state(const OracleEnv* e, float* dst) {
    memcpy(dst, e->state, sizeof(float) * (size_t)e->cfg.num_arenas * e->L.words);
}
void oracle_set_state(OracleEnv* e, const float* src) {
    memcpy(e->state, src, sizeof(float) * (size_t)e->cfg.num_arenas * e->L.words);
}
/* rasterise caller-supplied records (same contract as agar_observe_records) */
void oracle_observe_records(OracleEnv* e, const float* state, int n_records, float* obs) {
    int C = agar_spec_obs_channels(&e->cfg), G = e->cfg.grid_size, A = e->cfg.num_agents;
    size_t per = (size_t)A * C * G * G;
    for (int n = 0; n < n_records; ++n) observe_arena(e, state + (size_t)n * e->L.words, obs + (size_t)n * per);
}
void oracle_events(const OracleEnv* e, int32_t* events, int32_t* counts) {
    int cap = e->cfg.event_capacity;
    if (cap <= 0) return;
    memcpy(events, e->events, sizeof(int32_t) * (size_t)e->cfg.num_arenas * cap * 4);
    memcpy(counts, e->ev_count, sizeof(int32_t) * (size_t)e->cfg.num_arenas);
}
void oracle_config_default(AgarConfig* c) { agar_spec_config_default(c); }
void oracle_layout(const AgarConfig* c, AgarLayout* L) { agar_spec_layout(c, L); }
