/*
 * agar_b200.h — C-ABI of the B200-native batched Agar.io engine.
 *
 * This is the drop-in boundary for the env-step + grid-observation + GAE path of
 * jondeaton/AgarAI.  The reference has no FFI of its own on this path: its trainers
 * call a gym.Env duck type whose body lives in the external `gym_agario` package
 * (reference call sites: ppo/train.py:164,173  a2c/training.py:66,94
 * dqn/training.py:96,104).  Each entry point below cites the reference interface it
 * stands in for.  Plain pointers and sizes only; no torch / CUDA types in signatures
 * (streams are passed as `void*` holding a cudaStream_t; NULL = default stream).
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error; agar_last_error() gives
 *     the message (thread-local).  No exception crosses the ABI.
 *   - all `d_*` pointers are caller-owned DEVICE memory (e.g. torch CUDA tensors);
 *     `h_*` pointers are caller-owned HOST memory.  The library owns only the world
 *     state of the N arenas of a handle.
 *   - calls are asynchronous on the given stream unless the name ends in `_host`.
 *   - one AgarEnv per GPU; a handle is not thread-safe.
 */
#ifndef AGAR_B200_H
#define AGAR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AGAR_ABI_VERSION 1

/* ------------------------------------------------------------------------------------
 * Rule constants fixed by the spec.  Everything else is an AgarConfig field.
 * ---------------------------------------------------------------------------------- */
#define AGAR_MAX_CELLS 16 /* cell slots per player (split cap)                          */

/* observation channel flags, in output channel order
 * (features/extractors.py:46-64: pellets, own cells, others, viruses, foods)            */
#define AGAR_OBS_PELLETS 1
#define AGAR_OBS_CELLS 2
#define AGAR_OBS_OTHERS 4
#define AGAR_OBS_VIRUSES 8
#define AGAR_OBS_FOODS 16

/* act codes (ppo/action_conversion.py:47-60; a2c/train.py:76; dqn/training.py:225)      */
#define AGAR_ACT_NONE 0
#define AGAR_ACT_SPLIT 1
#define AGAR_ACT_FEED 2

/* Philox stream ids: counter = {slot, stream, tick, global_arena_id}, key = seed        */
#define AGAR_STREAM_PELLET_INIT 0
#define AGAR_STREAM_PELLET_RESPAWN 1
#define AGAR_STREAM_VIRUS_INIT 2
#define AGAR_STREAM_PLAYER_SPAWN 3
#define AGAR_STREAM_BOT_ROAM 4
#define AGAR_STREAM_VIRUS_RESPAWN 5

/* event types written to the optional per-step event log: int32[4] = {tick, type, a, b}  */
#define AGAR_EV_PELLET_EATEN 1 /* a = cell gid (player*16+slot), b = pellet index       */
#define AGAR_EV_FOOD_EATEN 2   /* a = cell gid, b = food slot                            */
#define AGAR_EV_VIRUS_POP 3    /* a = cell gid, b = virus index                          */
#define AGAR_EV_CELL_EATEN 4   /* a = eater gid, b = prey gid                            */
#define AGAR_EV_MERGE 5        /* a = surviving gid, b = absorbed gid                    */
#define AGAR_EV_SPLIT 6        /* a = source gid, b = new gid                            */
#define AGAR_EV_FEED 7         /* a = cell gid, b = food slot                            */
#define AGAR_EV_DEATH 8        /* a = player, b = 0                                      */
#define AGAR_EV_RESPAWN 9      /* a = player, b = 0                                      */
#define AGAR_EV_POP_PIECE 10   /* a = popped gid, b = new gid                            */

/* 16 unit directions used for virus-pop pieces (cos/sin(2*pi*k/16) rounded to f32;
 * literals so that host oracle and device agree bit for bit)                            */
#define AGAR_DIR16_X                                                                       \
    {1.0f, 0.92387953f, 0.70710678f, 0.38268343f, 0.0f, -0.38268343f, -0.70710678f,        \
     -0.92387953f, -1.0f, -0.92387953f, -0.70710678f, -0.38268343f, 0.0f, 0.38268343f,     \
     0.70710678f, 0.92387953f}
#define AGAR_DIR16_Y                                                                       \
    {0.0f, 0.38268343f, 0.70710678f, 0.92387953f, 1.0f, 0.92387953f, 0.70710678f,          \
     0.38268343f, 0.0f, -0.38268343f, -0.70710678f, -0.92387953f, -1.0f, -0.92387953f,     \
     -0.70710678f, -0.38268343f}

/* ------------------------------------------------------------------------------------
 * Configuration.  Stands in for gym.make(id, **kwargs):
 *   configuration/__init__.py:49-73 (num_agents, difficulty, ticks_per_step, arena_size,
 *   num_pellets, num_viruses, num_bots, pellet_regen, grid_size, num_frames, observe_*),
 *   a2c/train.py:17-64, dqn/train.py:55-81.
 * The rule constants below the kwargs are this build's own spec (the real engine is not
 * available: SURVEY.md §0.3) and are all overridable.
 * ---------------------------------------------------------------------------------- */
typedef struct AgarConfig {
    /* sizes */
    int32_t num_arenas;     /* N arenas owned by this handle (this GPU's shard)          */
    int32_t num_agents;     /* A  agents per arena (gym kw num_agents)                   */
    int32_t num_bots;       /* bots per arena (gym kw num_bots); players K = A + bots    */
    int32_t num_pellets;    /* P  (gym kw num_pellets)                                   */
    int32_t num_viruses;    /* V  (gym kw num_viruses)                                   */
    int32_t max_foods;      /* F  ejected-mass slots per arena                           */
    int32_t ticks_per_step; /* gym kw ticks_per_step (environment.proto:48, default 4)   */
    int32_t grid_size;      /* G  (gym kw grid_size, default 64)                         */
    int32_t obs_flags;      /* AGAR_OBS_* (gym kw observe_pellets/cells/others/viruses)  */
    int32_t pellet_regen;   /* gym kw pellet_regen                                       */
    int32_t auto_reset;     /* re-initialise an arena in the step in which all its agents
                               became done (vector-env semantics); 0 = dones are sticky  */
    int32_t event_capacity; /* events kept per arena per step; 0 disables the log        */
    int32_t bots_split;     /* reserved (bots never split/feed in this spec)             */
    int32_t reserved0;
    int64_t arena_id_base;  /* global id of local arena 0 (RNG is keyed by global id so
                               traces do not depend on how arenas are sharded over GPUs) */
    /* world */
    float arena_size;  /* gym kw arena_size                                              */
    float view_size;   /* side of the square window the grid observation covers          */
    /* masses */
    float start_mass;      /* 10: reference trainers assume it (a2c/training.py:257)     */
    float pellet_mass;     /* 1                                                          */
    float food_mass;       /* 10  mass ejected by act=2 and gained by eating it          */
    float virus_mass;      /* 100                                                        */
    float split_min_mass;  /* 20  a cell needs this much to split                        */
    float feed_min_mass;   /* 30  a cell needs this much to eject food                   */
    float virus_pop_mass;  /* 110 a cell at least this heavy pops on a virus             */
    float eat_ratio;       /* 1.25 eater mass >= eat_ratio * prey mass                   */
    /* kinematics (distance units per tick) */
    float r2_per_mass;     /* radius^2 = mass * r2_per_mass   (1/pi: area == mass)       */
    float speed_k;         /* max_speed = speed_k / sqrt(radius)                         */
    float target_reach;    /* agent target point = centroid + action * target_reach      */
    float split_speed;     /* launch impulse of the new half                             */
    float impulse_decay;   /* impulse *= impulse_decay each tick                         */
    float feed_speed;      /* launch speed of ejected food                               */
    float food_decay;      /* food velocity *= food_decay each tick                      */
    float food_spawn_gap;  /* food appears at radius + gap from the cell centre          */
    float pop_speed;       /* impulse of virus-pop pieces                                */
    float merge_delay;     /* ticks before a split/popped cell may merge again           */
    int32_t pop_pieces;    /* max pieces a popped cell throws off                        */
    /* bots */
    int32_t bot_roam_period; /* ticks between roam way-points (power of two)             */
    float bot_sense_radius;  /* bots chase the nearest pellet within this square window  */
    float mass_decay;        /* fraction of its mass a cell loses per tick while the rest stays >=
                                start_mass (applied at the start of the tick's motion); 0 = off */
} AgarConfig;

/* Offsets (in 32-bit words) of each field inside one arena's state record.  The world
 * state is `num_arenas` consecutive records of `words` words: a struct-of-arrays per arena,
 * arena-major, so that one CTA stages one record with bulk copies.  Every segment is
 * padded to a multiple of 4 words (16 B).  Integer fields are stored as int32 bit
 * patterns in the same buffer. */
typedef struct AgarLayout {
    int32_t words;                       /* record size                                  */
    int32_t pellet_x, pellet_y;          /* [P]  x < 0 marks an absent pellet            */
    int32_t virus_x, virus_y;            /* [V]                                          */
    int32_t dyn_begin;                   /* first word of the per-step rewritten part    */
    int32_t cell_x, cell_y;              /* [K*16] player-major, slot-minor              */
    int32_t cell_ix, cell_iy;            /* [K*16] launch impulse                        */
    int32_t cell_m;                      /* [K*16] mass, 0 = empty slot                  */
    int32_t cell_t;                      /* [K*16] merge timer in ticks (integral float) */
    int32_t food_x, food_y;              /* [F]  x < 0 marks a free slot                 */
    int32_t food_vx, food_vy;            /* [F]                                          */
    int32_t agent_done;                  /* [A]  int32 0/1                               */
    int32_t tick;                        /* [1]  uint32 ticks since agar_seed            */
    int32_t episode_steps;               /* [1]  int32 steps since the arena's reset     */
    int32_t dyn_end;
    int32_t num_players, cells_total;    /* K, K*16                                      */
    int32_t P_pad, V_pad, F_pad, A_pad;
} AgarLayout;

typedef struct AgarEnv AgarEnv;

/* thread-local message of the last failing call */
const char* agar_last_error(void);
int agar_abi_version(void);

/* fill *cfg with the default spec; the caller then overrides kwargs. */
int agar_config_default(AgarConfig* cfg);
/* pure function of the config; also used by the CPU oracle so both share one layout */
int agar_layout_compute(const AgarConfig* cfg, AgarLayout* out);
/* number of observation channels for cfg->obs_flags (features/extractors.py:30 `depth`) */
int agar_obs_channels(const AgarConfig* cfg);

/* gym.make(...)  — configuration/__init__.py:45-73, ppo/train.py:361-372 */
int agar_create(const AgarConfig* cfg, int device, AgarEnv** out);
int agar_destroy(AgarEnv* env);
int agar_get_config(const AgarEnv* env, AgarConfig* out);
int agar_get_layout(const AgarEnv* env, AgarLayout* out);

/* env.seed(): never called by the reference on the Agar.io env;
 * added so that rollouts are reproducible.  Resets the tick counters; call agar_reset
 * afterwards. */
int agar_seed(AgarEnv* env, uint64_t seed);

/* env.reset() — ppo/train.py:164, a2c/training.py:66, dqn/training.py:96.
 * d_reset_mask: uint8[N] (NULL = all arenas).  d_obs: float[N,A,C,G,G] or NULL. */
int agar_reset(AgarEnv* env, const uint8_t* d_reset_mask, float* d_obs, void* stream);

/* env.step(actions) — ppo/train.py:173, a2c/training.py:94, dqn/training.py:104.
 *   d_target_xy float[N,A,2]  direction*magnitude in the unit disc (action_conversion.py:31-44)
 *   d_act       int32[N,A]    AGAR_ACT_*
 *   d_obs       float[N,A,C,G,G] (NULL: no observation is rendered)
 *   d_reward    float[N,A]    change of the agent's total mass over the step
 *   d_done      uint8[N,A]
 * One launch advances every arena by ticks_per_step ticks and renders the observations. */
int agar_step(AgarEnv* env, const float* d_target_xy, const int32_t* d_act, float* d_obs,
              float* d_reward, uint8_t* d_done, void* stream);

/* Same, with the discrete action decode of ppo/action_conversion.py:19-80 (or the
 * unravel_index form of a2c/train.py:67-79, dqn/training.py:220-228) fused in:
 * action index -> (target, act) through a table of n_actions rows set beforehand. */
int agar_set_action_table(AgarEnv* env, const float* h_table_xy /*[n,2]*/,
                          const int32_t* h_table_act /*[n]*/, int32_t n_actions);
int agar_step_indexed(AgarEnv* env, const int32_t* d_action_index /*[N,A]*/, float* d_obs,
                      float* d_reward, uint8_t* d_done, void* stream);

/* GridFeatureExtractor.extract on the current state — features/extractors.py:39-96 —
 * for every agent of every arena, without stepping. */
int agar_observe(AgarEnv* env, float* d_obs, void* stream);

/* The same rasteriser on caller-supplied state records (e.g. snapshots kept in a rollout
 * buffer instead of 48 KB observations): d_state float[n_records, layout.words]. */
int agar_observe_records(AgarEnv* env, const float* d_state, int32_t n_records, float* d_obs,
                         void* stream);

/* whole-state copy out / in: float[N, layout.words] device buffers (parity tests,
 * checkpoint/resume). */
int agar_get_state(AgarEnv* env, float* d_dst, void* stream);
int agar_set_state(AgarEnv* env, const float* d_src, void* stream);

/* event log of the LAST step: d_events int32[N, event_capacity, 4], d_counts int32[N]
 * (count may exceed capacity; the excess was dropped).  Order inside a step is not
 * defined; compare as a multiset. */
int agar_events(AgarEnv* env, int32_t* d_events, int32_t* d_counts, void* stream);

/* Host-buffer step for callers that keep numpy arrays (the gym-compat view): copies the
 * actions host->device, steps, delivers obs/reward/done into the host buffers and synchronises.
 * h_obs may be NULL.  Host buffers should be pinned for full PCIe speed.  The call is ordered
 * after work enqueued earlier on caller streams through this handle.  For batches of 256 agents
 * or more the dense grids do not cross PCIe: each agent's centroid and occupied bins do (a few KB
 * instead of 48 KB at 3x64x64) and worker threads of the library rebuild the identical float32
 * grids in h_obs (AGAR_HOST_THREADS, default: the process's cores divided by LOCAL_WORLD_SIZE, less
 * one for the calling thread; the workers are pinned to this rank's share of the process's CPU set
 * and the calling thread stays on the free core of the share for the duration of the call -
 * AGAR_HOST_PIN=0 leaves all of them to the scheduler; AGAR_HOST_DENSE=1 restores the plain image
 * copy). */
int agar_step_host(AgarEnv* env, const float* h_target_xy, const int32_t* h_act, float* h_obs,
                   float* h_reward, uint8_t* h_done);

/* Persistent host observation buffer (off by default).  With on = 1 the caller promises that the
 * h_obs it passes to consecutive agar_step_host calls is the same memory and that nothing wrote to
 * it since the previous call returned (the usual vector-env arrangement: the env owns one
 * observation buffer, rollout/multi_rollout.py copies out of it).  The library then brings the
 * buffer from the previous step's grids to this step's by rewriting only what changed (the bins
 * that were or become occupied and the out-of-bounds rows / columns that moved: ~14 % of the
 * cache lines of a 3x64x64 grid per step) instead of streaming every grid out again.  The result
 * in h_obs is bit-identical to the default delivery.  A call with another pointer, the first
 * call, and agents whose image was copied whole fall back to the full rewrite by themselves. */
int agar_host_obs_persistent(AgarEnv* env, int on);

/* rl.gae — rl/__init__.py:94-137 (numpy statement rl/test.py:12-21) and
 * rl.make_returns — rl/__init__.py:140-159 (rl/test.py:24-30; a2c/losses.py:66-78) over a
 * time-major rollout buffer, one column per agent:
 *   d_reward float[T,N], d_value float[T+1,N], d_done uint8[T,N] or NULL,
 *   d_adv float[T,N] (may be NULL), d_ret float[T,N] (may be NULL),
 *   d_end_value float[N] or NULL (returns bootstrap; NULL = 0; the PPO caller passes
 *   value[T], ppo/train.py:273).
 * d_done == NULL reproduces the reference exactly (no masking inside; callers zero the
 * bootstrap, ppo/train.py:189-195).  With d_done, done[t]!=0 cuts the recursion after t.
 * gamma and lam are doubles because the reference multiplies them as Python floats
 * before they meet float32 data. */
int agar_gae(const float* d_reward, const float* d_value, const uint8_t* d_done,
             const float* d_end_value, double gamma, double lam, float* d_adv, float* d_ret,
             int32_t T, int32_t N, void* stream);

/* FeatureExtractor.extract — features/extractors.py:99-149 (helpers :171-212), the fixed-length "RAM-style"
 * observation the DQN trainer consumes (dqn/train.py:89-96): for every agent of every arena, from the current
 * state,  [nearest n_pellet pellets | n_virus viruses | n_food foods] as (x, y) relative to the agent's
 * centroid, the agent's own cells (absolute) and the n_other closest players' cells (relative), n_cell rows
 * of (x, y, impulse_x, impulse_y, mass) each; zero rows pad missing entities, -1000 rows missing players.
 * size = 2*(n_pellet+n_virus+n_food) + 5*(1+n_other)*n_cell (580 with the reference defaults 50,5,10,5,15).
 * d_out float[N, A, size].  An agent without cells gets zeros (the reference returns None).
 * Distance ties keep index order (numpy leaves them unspecified); `largest_cells` keeps the n_cell SMALLEST
 * cells, as the reference does (argsort ascending, :182-184). */
int agar_ram_feature_size(int32_t n_pellet, int32_t n_virus, int32_t n_food, int32_t n_other,
                          int32_t n_cell);
int agar_observe_ram(AgarEnv* env, int32_t n_pellet, int32_t n_virus, int32_t n_food, int32_t n_other,
                     int32_t n_cell, float* d_out, void* stream);

/* ScreenFeatureExtractor.extract — features/extractors.py:152-164: mean over the trailing colour axis of
 * uint8 frames, d_rgb uint8[n_pixels, 3] -> d_out float64[n_pixels] (out_f64 != 0: the reference's dtype,
 * bit-exact) or float32[n_pixels] (that value rounded once).  Needs no env handle. */
int agar_screen_gray(const uint8_t* d_rgb, int64_t n_pixels, void* d_out, int32_t out_f64, void* stream);

/* number of kernel launches issued by this handle since creation (bench bookkeeping) */
int64_t agar_launch_count(const AgarEnv* env);
/* bytes the last agar_step_host call moved host->device and device->host (bench bookkeeping: with the compact copy the
 * observations cross PCIe as per-agent lists of occupied bins, not as dense grids) */
int agar_host_copy_bytes(const AgarEnv* env, int64_t* h2d_bytes, int64_t* d2h_bytes);
/* which instantiation of the step kernel the handle launches: "generic" (sizes read at run time) or the name of a
 * compile-time specialised world ("cfg1", "cfg2", "cfg4", "a2c", "a2c_test", "dqn") */
const char* agar_kernel_variant(const AgarEnv* env);
/* dynamic shared memory one CTA of the step kernel needs for this config (bytes; -1 for a bad config) */
int64_t agar_step_smem_bytes(const AgarConfig* cfg);

#ifdef __cplusplus
}
#endif
#endif /* AGAR_B200_H */
