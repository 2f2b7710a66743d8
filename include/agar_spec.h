/*
 * agar_spec.h — the parts of the spec that are pure data: default constants and the state
 * record layout.  Header-only so the CUDA library and the CPU oracle (test infrastructure)
 * agree on the record format without linking one another.  No algorithm lives here.
 */
#ifndef AGAR_SPEC_H
#define AGAR_SPEC_H

#include <string.h>

#include "agar_b200.h"

static inline int agar_spec_pad4(int n) { return (n + 3) & ~3; }

static inline void agar_spec_config_default(AgarConfig* c) {
    memset(c, 0, sizeof(*c));
    c->num_arenas = 1;
    c->num_agents = 1;       /* environment.proto:61 default                              */
    c->num_bots = 0;
    c->num_pellets = 1000;   /* ppo.textproto:8                                           */
    c->num_viruses = 0;
    c->max_foods = 64;
    c->ticks_per_step = 4;   /* environment.proto:48                                      */
    c->grid_size = 64;       /* environment.proto:50                                      */
    c->obs_flags = AGAR_OBS_PELLETS | AGAR_OBS_CELLS | AGAR_OBS_OTHERS | AGAR_OBS_VIRUSES;
    c->pellet_regen = 1;
    c->auto_reset = 0;
    c->event_capacity = 0;
    c->arena_id_base = 0;
    c->arena_size = 1000.0f; /* ppo.textproto:9                                           */
    c->view_size = 128.0f;
    c->start_mass = 10.0f;
    c->pellet_mass = 1.0f;
    c->food_mass = 10.0f;
    c->virus_mass = 100.0f;
    c->split_min_mass = 20.0f;
    c->feed_min_mass = 30.0f;
    c->virus_pop_mass = 110.0f;
    c->eat_ratio = 1.25f;
    c->r2_per_mass = 0.31830987f; /* 1/pi */
    c->speed_k = 3.0f;
    c->target_reach = 64.0f;
    c->split_speed = 6.0f;
    c->impulse_decay = 0.9f;
    c->feed_speed = 8.0f;
    c->food_decay = 0.85f;
    c->food_spawn_gap = 1.0f;
    c->pop_speed = 5.0f;
    c->merge_delay = 400.0f;
    c->pop_pieces = 7;
    c->bot_roam_period = 256;
    c->bot_sense_radius = 64.0f;
}

static inline int agar_spec_obs_channels(const AgarConfig* c) {
    int n = 0, f = c->obs_flags;
    for (int b = 0; b < 5; ++b) n += (f >> b) & 1;
    return n;
}

/* returns 0 if the config is usable, else a static message index > 0 */
static inline const char* agar_spec_config_check(const AgarConfig* c) {
    if (c->num_arenas < 1) return "num_arenas must be >= 1";
    if (c->num_agents < 1) return "num_agents must be >= 1";
    if (c->num_bots < 0) return "num_bots must be >= 0";
    if (c->num_agents + c->num_bots > 64) return "num_agents + num_bots must be <= 64";
    if (c->num_pellets < 0 || c->num_pellets > 8192) return "num_pellets must be in [0, 8192]";
    if (c->num_viruses < 0 || c->num_viruses > 256) return "num_viruses must be in [0, 256]";
    if (c->max_foods < 0 || c->max_foods > 256) return "max_foods must be in [0, 256]";
    if (c->ticks_per_step < 1 || c->ticks_per_step > 64) return "ticks_per_step must be in [1, 64]";
    if (c->grid_size < 4 || c->grid_size > 128 || (c->grid_size & 3))
        return "grid_size must be a multiple of 4 in [4, 128]";
    if (!(c->obs_flags & 31)) return "obs_flags selects no channel";
    if (!(c->arena_size > 0.0f)) return "arena_size must be > 0";
    if (!(c->view_size > 0.0f)) return "view_size must be > 0";
    if (!(c->target_reach > 0.0f)) return "target_reach must be > 0";
    if (!(c->start_mass > 0.0f)) return "start_mass must be > 0";
    if (!(c->r2_per_mass > 0.0f)) return "r2_per_mass must be > 0";
    if (!(c->split_min_mass > 0.0f)) return "split_min_mass must be > 0";
    if (!(c->feed_min_mass > c->food_mass)) return "feed_min_mass must exceed food_mass";
    if (!(c->virus_pop_mass > 0.0f)) return "virus_pop_mass must be > 0";
    if (c->pop_pieces < 0 || c->pop_pieces > 15) return "pop_pieces must be in [0, 15]";
    if (c->bot_roam_period < 1 || (c->bot_roam_period & (c->bot_roam_period - 1)))
        return "bot_roam_period must be a power of two";
    if (!(c->mass_decay >= 0.0f && c->mass_decay < 0.5f)) return "mass_decay must be in [0, 0.5)";
    if (c->event_capacity < 0 || c->event_capacity > 65536) return "event_capacity out of range";
    return 0;
}

static inline void agar_spec_layout(const AgarConfig* c, AgarLayout* L) {
    int K = c->num_agents + c->num_bots;
    int KC = K * AGAR_MAX_CELLS;
    int Pp = agar_spec_pad4(c->num_pellets), Vp = agar_spec_pad4(c->num_viruses);
    int Fp = agar_spec_pad4(c->max_foods), Ap = agar_spec_pad4(c->num_agents);
    int o = 0;
    memset(L, 0, sizeof(*L));
    L->num_players = K; L->cells_total = KC;
    L->P_pad = Pp; L->V_pad = Vp; L->F_pad = Fp; L->A_pad = Ap;
    L->pellet_x = o; o += Pp;
    L->pellet_y = o; o += Pp;
    L->virus_x = o; o += Vp;
    L->virus_y = o; o += Vp;
    L->dyn_begin = o;
    L->cell_x = o; o += KC;
    L->cell_y = o; o += KC;
    L->cell_ix = o; o += KC;
    L->cell_iy = o; o += KC;
    L->cell_m = o; o += KC;
    L->cell_t = o; o += KC;
    L->food_x = o; o += Fp;
    L->food_y = o; o += Fp;
    L->food_vx = o; o += Fp;
    L->food_vy = o; o += Fp;
    L->agent_done = o; o += Ap;
    L->tick = o; L->episode_steps = o + 1; o += 4;
    L->dyn_end = o;
    L->words = o;
}

#endif /* AGAR_SPEC_H */
