/*
 * mapf_oracle.c -- CPU oracle (TEST INFRASTRUCTURE ONLY, see mapf_oracle.h).
 *
 * A literal, scalar restatement of the reference's MAPF transition.  "ENV:a-b" cites
 * /root/reference/src/environments/reference_model_multi_agent.py lines a-b.
 * Coordinates are (row, col) like the reference's (y, x).
 */
#include "mapf_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define UNASSIGNED (-1) /* ENV:32 */

struct oracle_env {
    oracle_config cfg;
    int R, C, N, V, H, F;
    uint8_t *grid;     /* [R*C] 0 free, 1 obstacle */
    int16_t *free_pos; /* [F,2] ENV:82 */
    int16_t *starts, *positions, *goals; /* [N,2] ENV:83-85 */
    uint8_t *reached, *completed_once;   /* ENV:86-87 */
    float *bp_prev;                      /* ENV:89 */
    int16_t *occ_owner, *goal_owner;     /* [R*C] ENV:102-103 */
    uint8_t *h_gp, *h_mv, *h_fm;         /* [H,N] ENV:115-117 */
    int16_t *h_dist;                     /* [H,N] ENV:118 */
    int hist_count, hist_head;           /* ENV:119-120 */
    int step_count;
    double ep_goals_total, ep_blocking;  /* ENV:88, ENV:63 */
    double ep_dl_events, ep_ll_events, ep_dl_steps, ep_ll_steps; /* ENV:64-67 */
    int dl_prev, ll_prev;                /* ENV:68-69 */
    uint64_t rng;
    /* device-RNG replay (oracle_set_philox): the draws of the CUDA kernels, restated */
    int ph_on;
    uint64_t ph_seed;
    int64_t ph_stream;
    uint32_t ph_counter;
    /* scratch, ENV:90-101 */
    int16_t *prev_pos, *intended;
    uint8_t *reached_goal, *moved, *failed, *gprog, *prev_on_goal, *cur_on_goal;
    int16_t *dist;
    float *goal_step_flags, *blocking_flags;
    int8_t *actions_taken;
    int32_t *last_cand;
    int32_t *last_rank; /* rank drawn by the last reassignment of each agent (debug / replay) */
    int32_t *perm; /* [F] scratch for sampling without replacement */
    uint8_t *tmp_obs; /* [V*V] */
};

/* ---------------------------------------------------------------- RNG (oracle-own stream) */
static uint64_t splitmix64(uint64_t *s) {
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static uint32_t rng_below(uint64_t *s, uint32_t n) { /* unbiased, rejection */
    uint32_t lim = (uint32_t)(0x100000000ull - (0x100000000ull % n));
    for (;;) {
        uint32_t x = (uint32_t)(splitmix64(s) >> 32);
        if (lim == 0 || x < lim) return x % n;
    }
}

/* ---------------------------------------------------------------- device-RNG replay
 * The reference draws with numpy PCG64 (ENV:76-78,277,300); the CUDA kernels draw with
 * Philox4x32-10 keyed by (seed, global env id).  North-star accepts a different stream, so this
 * part has no reference counterpart: it restates the KERNELS' documented draw discipline
 * (include/mapf_b200.h "RNG") so that the benchmarked mode -- Philox goal draws, in-launch
 * auto-reset, fused masked sampler -- can be compared bit for bit over whole episodes. */
static void philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c[4]) {
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
void oracle_philox_raw(uint32_t k0, uint32_t k1, uint32_t ctr[4]) { philox4x32_10(k0, k1, ctr); }
/* key = seed ^ golden * (stream + 1), stream = global env id */
static void philox_env(uint64_t seed, int64_t stream, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                       uint32_t out[4]) {
    uint64_t k = seed ^ (0x9E3779B97F4A7C15ull * (uint64_t)(stream + 1));
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
    philox4x32_10((uint32_t)k, (uint32_t)(k >> 32), out);
}
static uint32_t mulhi32(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }

void oracle_set_philox(oracle_env *e, int on, uint64_t seed, int64_t env_global, uint32_t counter) {
    e->ph_on = on; e->ph_seed = seed; e->ph_stream = env_global; e->ph_counter = counter;
}
uint32_t oracle_get_philox_counter(const oracle_env *e) { return e->ph_counter; }

/* the benchmark's samplers (scripts/benchmark_multi_agent_env.py:38-57) with the kernels' draws: one
 * Philox block per agent quad, word a & 3; masked = uniform over the set bits of the 5-entry mask */
void oracle_sample_actions_philox(const oracle_env *e, uint64_t call_counter, int masked,
                                  const int8_t *mask /* [N,5] */, int8_t *actions /* [N] */) {
    for (int a = 0; a < e->N; ++a) {
        uint32_t x4[4];
        philox_env(e->ph_seed ^ 0xA511E9B3ull, e->ph_stream, (uint32_t)call_counter,
                   (uint32_t)(call_counter >> 32), (uint32_t)(a >> 2), 0x41435421u, x4);
        uint32_t x = x4[a & 3];
        if (!masked) { actions[a] = (int8_t)mulhi32(x, 5u); continue; }
        int valid[5], nv = 0;
        for (int k = 0; k < 5; ++k) if (mask[5 * a + k]) valid[nv++] = k;
        actions[a] = nv ? (int8_t)valid[mulhi32(x, (uint32_t)nv)] : 0;
    }
}

/* ---------------------------------------------------------------- construction */
static void *zalloc(size_t n) { return calloc(n ? n : 1, 1); }

oracle_env *oracle_create(const oracle_config *cfg, const uint8_t *grid, uint64_t seed) {
    if (!cfg || !grid || cfg->rows <= 0 || cfg->cols <= 0 || cfg->num_agents <= 0) return NULL;
    oracle_env *e = (oracle_env *)zalloc(sizeof(*e));
    e->cfg = *cfg;
    /* ENV:56-60 clamp to >= 1 */
    if (e->cfg.deadlock_window_steps < 1) e->cfg.deadlock_window_steps = 1;
    if (e->cfg.livelock_window_steps < 1) e->cfg.livelock_window_steps = 1;
    if (e->cfg.lock_nearby_manhattan < 1) e->cfg.lock_nearby_manhattan = 1;
    if (e->cfg.lock_min_neighbors < 1) e->cfg.lock_min_neighbors = 1;
    e->R = cfg->rows; e->C = cfg->cols; e->N = cfg->num_agents;
    e->V = 2 * cfg->sensor_range + 1; /* ENV:137 */
    e->H = e->cfg.deadlock_window_steps > e->cfg.livelock_window_steps
               ? e->cfg.deadlock_window_steps : e->cfg.livelock_window_steps; /* ENV:114 */
    int RC = e->R * e->C, N = e->N;
    e->grid = (uint8_t *)zalloc(RC);
    memcpy(e->grid, grid, RC);
    e->free_pos = (int16_t *)zalloc(sizeof(int16_t) * 2 * RC);
    e->F = 0;
    for (int r = 0; r < e->R; ++r)
        for (int c = 0; c < e->C; ++c)
            if (grid[r * e->C + c] == 0) { /* ENV:82 argwhere(grid == EMPTY) */
                e->free_pos[2 * e->F] = (int16_t)r;
                e->free_pos[2 * e->F + 1] = (int16_t)c;
                e->F++;
            }
    e->starts = (int16_t *)zalloc(sizeof(int16_t) * 2 * N);
    e->positions = (int16_t *)zalloc(sizeof(int16_t) * 2 * N);
    e->goals = (int16_t *)zalloc(sizeof(int16_t) * 2 * N);
    e->reached = (uint8_t *)zalloc(N);
    e->completed_once = (uint8_t *)zalloc(N);
    e->bp_prev = (float *)zalloc(sizeof(float) * N);
    e->occ_owner = (int16_t *)zalloc(sizeof(int16_t) * RC);
    e->goal_owner = (int16_t *)zalloc(sizeof(int16_t) * RC);
    e->h_gp = (uint8_t *)zalloc((size_t)e->H * N);
    e->h_mv = (uint8_t *)zalloc((size_t)e->H * N);
    e->h_fm = (uint8_t *)zalloc((size_t)e->H * N);
    e->h_dist = (int16_t *)zalloc(sizeof(int16_t) * (size_t)e->H * N);
    e->prev_pos = (int16_t *)zalloc(sizeof(int16_t) * 2 * N);
    e->intended = (int16_t *)zalloc(sizeof(int16_t) * 2 * N);
    e->reached_goal = (uint8_t *)zalloc(N);
    e->moved = (uint8_t *)zalloc(N);
    e->failed = (uint8_t *)zalloc(N);
    e->gprog = (uint8_t *)zalloc(N);
    e->prev_on_goal = (uint8_t *)zalloc(N);
    e->cur_on_goal = (uint8_t *)zalloc(N);
    e->dist = (int16_t *)zalloc(sizeof(int16_t) * N);
    e->goal_step_flags = (float *)zalloc(sizeof(float) * N);
    e->blocking_flags = (float *)zalloc(sizeof(float) * N);
    e->actions_taken = (int8_t *)zalloc(N);
    e->last_cand = (int32_t *)zalloc(sizeof(int32_t) * N);
    e->last_rank = (int32_t *)zalloc(sizeof(int32_t) * N);
    e->perm = (int32_t *)zalloc(sizeof(int32_t) * RC);
    e->tmp_obs = (uint8_t *)zalloc((size_t)e->V * e->V);
    e->rng = seed * 0x9E3779B97F4A7C15ull + 0x1234567ull;
    for (int i = 0; i < RC; ++i) { e->occ_owner[i] = UNASSIGNED; e->goal_owner[i] = UNASSIGNED; }
    return e;
}

void oracle_destroy(oracle_env *e) {
    if (!e) return;
    free(e->grid); free(e->free_pos); free(e->starts); free(e->positions); free(e->goals);
    free(e->reached); free(e->completed_once); free(e->bp_prev); free(e->occ_owner);
    free(e->goal_owner); free(e->h_gp); free(e->h_mv); free(e->h_fm); free(e->h_dist);
    free(e->prev_pos); free(e->intended); free(e->reached_goal); free(e->moved); free(e->failed);
    free(e->gprog); free(e->prev_on_goal); free(e->cur_on_goal); free(e->dist);
    free(e->goal_step_flags); free(e->blocking_flags); free(e->actions_taken); free(e->last_cand); free(e->last_rank);
    free(e->perm); free(e->tmp_obs);
    free(e);
}

int oracle_num_free(const oracle_env *e) { return e->F; }
void oracle_get_free_positions(const oracle_env *e, int16_t *out) {
    memcpy(out, e->free_pos, sizeof(int16_t) * 2 * e->F);
}

/* ENV:200-205 */
static void rebuild_occupancy_owner(oracle_env *e) {
    for (int i = 0; i < e->R * e->C; ++i) e->occ_owner[i] = UNASSIGNED;
    for (int a = 0; a < e->N; ++a)
        e->occ_owner[e->positions[2 * a] * e->C + e->positions[2 * a + 1]] = (int16_t)a;
}
/* ENV:207-212 */
static void rebuild_goal_owner(oracle_env *e) {
    for (int i = 0; i < e->R * e->C; ++i) e->goal_owner[i] = UNASSIGNED;
    for (int a = 0; a < e->N; ++a)
        e->goal_owner[e->goals[2 * a] * e->C + e->goals[2 * a + 1]] = (int16_t)a;
}

static int in_grid(const oracle_env *e, int r, int c) {
    return r >= 0 && r < e->R && c >= 0 && c < e->C;
}

int oracle_set_layout(oracle_env *e, const int16_t *starts, const int16_t *goals) {
    for (int a = 0; a < e->N; ++a) {
        if (!in_grid(e, starts[2 * a], starts[2 * a + 1])) return ORACLE_ERR_BAD_ARG;
        if (!in_grid(e, goals[2 * a], goals[2 * a + 1])) return ORACLE_ERR_BAD_ARG;
    }
    memcpy(e->starts, starts, sizeof(int16_t) * 2 * e->N);
    memcpy(e->positions, starts, sizeof(int16_t) * 2 * e->N); /* ENV:130 / ENV:279 */
    memcpy(e->goals, goals, sizeof(int16_t) * 2 * e->N);
    rebuild_goal_owner(e);
    rebuild_occupancy_owner(e);
    return ORACLE_OK;
}

void oracle_reset_lock_tracking(oracle_env *e) { /* ENV:360-372 */
    size_t hn = (size_t)e->H * e->N;
    memset(e->h_gp, 0, hn); memset(e->h_mv, 0, hn); memset(e->h_fm, 0, hn);
    memset(e->h_dist, 0, sizeof(int16_t) * hn);
    e->hist_count = 0; e->hist_head = 0;
    e->ep_dl_events = e->ep_ll_events = e->ep_dl_steps = e->ep_ll_steps = 0.0;
    e->dl_prev = e->ll_prev = 0;
}

int oracle_set_state(oracle_env *e, const int16_t *positions, const int16_t *starts,
                     const int16_t *goals, const uint8_t *reached, const uint8_t *completed_once,
                     const float *blocking_prev, const int32_t *step_count,
                     const double *episode_goals_total) {
    int N = e->N;
    if (positions) memcpy(e->positions, positions, sizeof(int16_t) * 2 * N);
    if (starts) memcpy(e->starts, starts, sizeof(int16_t) * 2 * N);
    if (goals) memcpy(e->goals, goals, sizeof(int16_t) * 2 * N);
    if (reached) memcpy(e->reached, reached, N);
    if (completed_once) memcpy(e->completed_once, completed_once, N);
    if (blocking_prev) memcpy(e->bp_prev, blocking_prev, sizeof(float) * N);
    if (step_count) e->step_count = *step_count;
    if (episode_goals_total) e->ep_goals_total = *episode_goals_total;
    for (int a = 0; a < N; ++a) {
        if (!in_grid(e, e->positions[2 * a], e->positions[2 * a + 1])) return ORACLE_ERR_BAD_ARG;
        if (!in_grid(e, e->goals[2 * a], e->goals[2 * a + 1])) return ORACLE_ERR_BAD_ARG;
    }
    rebuild_goal_owner(e);
    rebuild_occupancy_owner(e);
    return ORACLE_OK;
}

/* ---------------------------------------------------------------- observation channels */
/* ENV:707-747 */
static void get_obs(const oracle_env *e, int idx, uint8_t *obs) {
    int V = e->V, sr = e->cfg.sensor_range;
    int base_r = e->positions[2 * idx] - sr;
    int base_c = e->positions[2 * idx + 1] - sr;
    for (int k = 0; k < V * V; ++k) obs[k] = 1; /* np.full(OBSTACLE_CELL) */
    for (int i = 0; i < V; ++i) {
        int r = base_r + i;
        if (r < 0 || r >= e->R) continue;
        for (int j = 0; j < V; ++j) {
            int c = base_c + j;
            if (c < 0 || c >= e->C) continue;
            if (e->grid[r * e->C + c] == 1) { obs[i * V + j] = 1; continue; }
            int occ = e->occ_owner[r * e->C + c];
            if (occ != UNASSIGNED && occ != idx) { obs[i * V + j] = 2; continue; }
            int g = e->goal_owner[r * e->C + c];
            if (g == idx) obs[i * V + j] = 3;
            else if (g != UNASSIGNED) obs[i * V + j] = 4;
            else obs[i * V + j] = 0;
        }
    }
}

static int traversable(uint8_t v) { return v == 0 || v == 3 || v == 4; } /* ENV:31 */

/* ENV:749-773 */
static void get_action_mask(const oracle_env *e, const uint8_t *obs, int8_t *mask) {
    int V = e->V, x = e->cfg.sensor_range, y = e->cfg.sensor_range;
    mask[0] = 1; mask[1] = mask[2] = mask[3] = mask[4] = 0;
    if (x > 0 && traversable(obs[(x - 1) * V + y])) mask[1] = 1;
    if (y < V - 1 && traversable(obs[x * V + y + 1])) mask[2] = 1;
    if (x < V - 1 && traversable(obs[(x + 1) * V + y])) mask[3] = 1;
    if (y > 0 && traversable(obs[x * V + y - 1])) mask[4] = 1;
}

/* ENV:330-335 with the denominators of ENV:152-155 */
static void get_goal_delta(const oracle_env *e, int idx, float *gd) {
    float d0 = (float)(int16_t)(e->goals[2 * idx] - e->positions[2 * idx]);
    float d1 = (float)(int16_t)(e->goals[2 * idx + 1] - e->positions[2 * idx + 1]);
    if (e->cfg.normalize_goal_delta) {
        float den0 = (float)(e->R - 1 > 1 ? e->R - 1 : 1);
        float den1 = (float)(e->C - 1 > 1 ? e->C - 1 : 1);
        d0 = d0 / den0;
        d1 = d1 / den1;
    }
    gd[0] = d0; gd[1] = d1;
}

/* everything _flatten_observation (ENV:306-328) needs for agent idx, from current state */
static void emit_agent_obs(oracle_env *e, int idx, oracle_outputs *out) {
    if (!out) return;
    int V2 = e->V * e->V;
    uint8_t *obs = out->local_obs ? out->local_obs + (size_t)idx * V2 : e->tmp_obs;
    get_obs(e, idx, obs);
    if (out->action_mask) get_action_mask(e, obs, out->action_mask + 5 * idx);
    float gd[2];
    get_goal_delta(e, idx, gd);
    if (out->goal_delta) { out->goal_delta[2 * idx] = gd[0]; out->goal_delta[2 * idx + 1] = gd[1]; }
    if (out->goal_distance) {
        float a0 = gd[0] < 0 ? -gd[0] : gd[0], a1 = gd[1] < 0 ? -gd[1] : gd[1];
        out->goal_distance[idx] = a0 + a1; /* ENV:320, float32 sum */
    }
    if (out->blocking_prev) out->blocking_prev[idx] = e->bp_prev[idx]; /* ENV:322 */
}

/* ---------------------------------------------------------------- layouts */
static int draw_layout(oracle_env *e) { /* ENV:267-282 with the oracle's own RNG */
    int need = 2 * e->N;
    if (e->F < need) return ORACLE_ERR_TOO_FEW_CELLS;
    if (e->ph_on) {
        /* Symmetric rejection, the kernels' rule: every slot (starts 0..N-1, then goals) draws a uniform free
         * cell; a slot that equals a lower-numbered slot redraws in the next round.  Round t of this env uses
         * Philox counter (ph_counter + t, agent, "RESE", 0): word 0 -> start, word 1 -> goal. */
        int N = e->N;
        int *cs = e->perm, *cg = e->perm + N;
        uint32_t rs = N >= 32 ? 0xFFFFFFFFu : ((1u << N) - 1u), rg = rs, rounds = 0;
        while (rs | rg) {
            for (int a = 0; a < N; ++a) {
                if (!(((rs | rg) >> a) & 1u)) continue;
                uint32_t x[4];
                philox_env(e->ph_seed, e->ph_stream, e->ph_counter + rounds, (uint32_t)a, 0x52455345u, 0, x);
                if ((rs >> a) & 1u) cs[a] = (int)mulhi32(x[0], (uint32_t)e->F);
                if ((rg >> a) & 1u) cg[a] = (int)mulhi32(x[1], (uint32_t)e->F);
            }
            rs = 0; rg = 0;
            for (int g = 0; g < N; ++g)
                for (int a = 0; a < N; ++a) {
                    if (a < g && cs[a] == cs[g]) rs |= 1u << g;
                    if (cs[a] == cg[g]) rg |= 1u << g;
                    if (a < g && cg[a] == cg[g]) rg |= 1u << g;
                }
            rounds++;
        }
        e->ph_counter += rounds;
        for (int a = 0; a < N; ++a) {
            e->starts[2 * a] = e->free_pos[2 * cs[a]]; e->starts[2 * a + 1] = e->free_pos[2 * cs[a] + 1];
            e->goals[2 * a] = e->free_pos[2 * cg[a]];  e->goals[2 * a + 1] = e->free_pos[2 * cg[a] + 1];
        }
        memcpy(e->positions, e->starts, sizeof(int16_t) * 2 * N);
        rebuild_goal_owner(e);
        rebuild_occupancy_owner(e);
        return ORACLE_OK;
    }
    for (int i = 0; i < e->F; ++i) e->perm[i] = i;
    for (int i = 0; i < need; ++i) { /* partial Fisher-Yates: uniform without replacement */
        int j = i + (int)rng_below(&e->rng, (uint32_t)(e->F - i));
        int t = e->perm[i]; e->perm[i] = e->perm[j]; e->perm[j] = t;
    }
    for (int a = 0; a < e->N; ++a) {
        e->starts[2 * a] = e->free_pos[2 * e->perm[a]];
        e->starts[2 * a + 1] = e->free_pos[2 * e->perm[a] + 1];
        e->goals[2 * a] = e->free_pos[2 * e->perm[e->N + a]];
        e->goals[2 * a + 1] = e->free_pos[2 * e->perm[e->N + a] + 1];
    }
    memcpy(e->positions, e->starts, sizeof(int16_t) * 2 * e->N);
    rebuild_goal_owner(e);
    rebuild_occupancy_owner(e);
    return ORACLE_OK;
}

/* ---------------------------------------------------------------- reset, ENV:440-472 */
int oracle_reset(oracle_env *e, int mode, const int16_t *starts, const int16_t *goals,
                 oracle_outputs *out) {
    e->step_count = 0;
    e->ep_blocking = 0.0;
    e->ep_goals_total = 0.0;
    oracle_reset_lock_tracking(e);
    memset(e->reached, 0, e->N);
    memset(e->completed_once, 0, e->N);
    for (int a = 0; a < e->N; ++a) e->bp_prev[a] = 0.0f;
    if (mode == 0) { /* ENV:452-455: positions only (F7) */
        memcpy(e->positions, e->starts, sizeof(int16_t) * 2 * e->N);
        rebuild_goal_owner(e);
        rebuild_occupancy_owner(e);
    } else if (mode == 1) {
        if (e->F < 2 * e->N) return ORACLE_ERR_TOO_FEW_CELLS;
        if (!starts || !goals) return ORACLE_ERR_BAD_ARG;
        int rc = oracle_set_layout(e, starts, goals);
        if (rc) return rc;
    } else {
        int rc = draw_layout(e);
        if (rc) return rc;
    }
    for (int a = 0; a < e->N; ++a) emit_agent_obs(e, a, out); /* ENV:459-468, final state */
    if (out) {
        if (out->reward) memset(out->reward, 0, sizeof(float) * e->N);
        if (out->terminated) *out->terminated = 0;
        if (out->truncated) *out->truncated = 0;
    }
    return ORACLE_OK;
}

/* ---------------------------------------------------------------- lifelong goals, ENV:284-304 */
static int assign_new_goal(oracle_env *e, int idx, int rank, const int16_t *override_rc) {
    int og_r = e->goals[2 * idx], og_c = e->goals[2 * idx + 1];
    e->goal_owner[og_r * e->C + og_c] = UNASSIGNED; /* ENV:288 */
    int n = 0;
    for (int f = 0; f < e->F; ++f) { /* ENV:290-295 candidates in free-cell order */
        int cell = e->free_pos[2 * f] * e->C + e->free_pos[2 * f + 1];
        if (e->occ_owner[cell] == UNASSIGNED && e->goal_owner[cell] == UNASSIGNED) e->perm[n++] = f;
    }
    e->last_cand[idx] = n;
    int r, c;
    if (override_rc && override_rc[0] >= 0) {
        r = override_rc[0]; c = override_rc[1];
        if (!in_grid(e, r, c)) return ORACLE_ERR_BAD_ARG;
    } else {
        if (n == 0) return ORACLE_ERR_NO_GOAL_CELL; /* ENV:296-298 */
        int k;
        if (rank >= 0) k = rank;
        else if (e->ph_on) { /* the kernels' draw: counter (ph_counter, agent, "GOAL", 0), word 0 scaled to [0, n) */
            uint32_t x[4];
            philox_env(e->ph_seed, e->ph_stream, e->ph_counter, (uint32_t)idx, 0x474F414Cu, 0, x);
            k = (int)mulhi32(x[0], (uint32_t)n);
            e->ph_counter++;
        } else k = (int)rng_below(&e->rng, (uint32_t)n); /* ENV:300 */
        if (k >= n) return ORACLE_ERR_BAD_ARG;
        e->last_rank[idx] = k;
        r = e->free_pos[2 * e->perm[k]];
        c = e->free_pos[2 * e->perm[k] + 1];
    }
    e->goals[2 * idx] = (int16_t)r; e->goals[2 * idx + 1] = (int16_t)c; /* ENV:302 */
    e->goal_owner[r * e->C + c] = (int16_t)idx;                           /* ENV:303 */
    return ORACLE_OK;
}

/* ---------------------------------------------------------------- lock metrics */
/* ENV:374-387 */
static void append_lock_history(oracle_env *e) {
    int row = e->hist_head, N = e->N;
    for (int a = 0; a < N; ++a) {
        e->h_gp[row * N + a] = e->gprog[a];
        e->h_mv[row * N + a] = e->moved[a];
        e->h_fm[row * N + a] = e->failed[a];
        e->h_dist[row * N + a] = e->dist[a];
    }
    e->hist_head = (e->hist_head + 1) % e->H;
    e->hist_count = e->hist_count + 1 < e->H ? e->hist_count + 1 : e->H;
}

/* ENV:400-438 (participants from ENV:389-398).  part[] scratch holds member indices. */
static void detect_lock_step(oracle_env *e, int *deadlock, int *livelock) {
    int N = e->N, H = e->H;
    *deadlock = 0; *livelock = 0;
    int any_off = 0;
    for (int a = 0; a < N; ++a) any_off |= !e->cur_on_goal[a];
    if (!any_off) return; /* ENV:401-402 */

    int dw = e->cfg.deadlock_window_steps, lw = e->cfg.livelock_window_steps;
    int near = e->cfg.lock_nearby_manhattan, minnb = e->cfg.lock_min_neighbors;
    /* two passes over the focal list, deadlock first (ENV:408-420) then livelock (ENV:422-436) */
    for (int pass = 0; pass < 2; ++pass) {
        int W = pass == 0 ? dw : lw;
        if (e->hist_count < W) continue;
        for (int f = 0; f < N; ++f) {
            if (e->cur_on_goal[f]) continue; /* ENV:391 focal = off-goal agents */
            int members[64]; /* N <= 32 enforced by callers; generous */
            int m = 0, nb = 0;
            members[m++] = f;
            for (int a = 0; a < N; ++a) {
                int d = abs(e->positions[2 * a] - e->positions[2 * f]) +
                        abs(e->positions[2 * a + 1] - e->positions[2 * f + 1]);
                if (d <= near && d > 0) { if (m < 64) members[m++] = a; nb++; } /* ENV:393-394 */
            }
            if (nb < minnb) continue; /* ENV:395-396 */
            double gp_sum = 0, mv_sum = 0, fm_sum = 0, d_start = 0, d_end = 0;
            for (int k = W; k >= 1; --k) { /* idxs = (head - arange(W,0,-1)) % H */
                int row = ((e->hist_head - k) % H + H) % H;
                for (int q = 0; q < m; ++q) {
                    int a = members[q];
                    gp_sum += e->h_gp[row * N + a];
                    mv_sum += e->h_mv[row * N + a];
                    fm_sum += e->h_fm[row * N + a];
                    if (k == W) d_start += e->h_dist[row * N + a]; /* window[0]  ENV:432 */
                    if (k == 1) d_end += e->h_dist[row * N + a];   /* window[-1] ENV:433 */
                }
            }
            if (pass == 0) {
                if (gp_sum <= 0.0 && mv_sum <= 0.0 && fm_sum > 0.0) { *deadlock = 1; return; }
            } else {
                double red = d_start - d_end;
                if (gp_sum <= 0.0 && mv_sum > 0.0 && red <= e->cfg.lock_progress_epsilon) {
                    *livelock = 1; return;
                }
            }
        }
    }
}

/* ---------------------------------------------------------------- step, ENV:474-695 */
int oracle_step(oracle_env *e, const int8_t *actions, const int32_t *goal_rank,
                const int16_t *goal_override, oracle_outputs *out) {
    int N = e->N, C = e->C;
    static const int dR[5] = {0, -1, 0, 1, 0}; /* ENV:104-113: no-op, up, right, down, left */
    static const int dC[5] = {0, 0, 1, 0, -1};
    float reward[64];
    if (N > 64) return ORACLE_ERR_BAD_ARG;

    e->step_count += 1; /* ENV:475 */
    int goal_reassigned = 0;
    for (int a = 0; a < N; ++a) {
        reward[a] = 0.0f;
        e->last_rank[a] = -1;
        e->reached_goal[a] = 0;
        e->goal_step_flags[a] = 0.0f;
        e->blocking_flags[a] = 0.0f;
        e->actions_taken[a] = 0;
    }
    memcpy(e->prev_pos, e->positions, sizeof(int16_t) * 2 * N); /* ENV:484 */

    for (int idx = 0; idx < N; ++idx) { /* ENV:502 sequential, order matters */
        int action = actions ? actions[idx] : 0;
        if (action < 0 || action > 4) return ORACLE_ERR_INVALID_ACTION; /* ENV:504-506 */
        e->actions_taken[idx] = (int8_t)action;
        int pr = e->positions[2 * idx], pc = e->positions[2 * idx + 1];
        int nr = pr + dR[action], nc = pc + dC[action];
        e->intended[2 * idx] = (int16_t)nr; e->intended[2 * idx + 1] = (int16_t)nc; /* ENV:514 */
        int valid = in_grid(e, nr, nc) && e->grid[nr * C + nc] == 0 &&
                    (e->occ_owner[nr * C + nc] == UNASSIGNED || e->occ_owner[nr * C + nc] == idx);
        if (valid && (nr != pr || nc != pc)) { /* ENV:522-526 */
            e->occ_owner[pr * C + pc] = UNASSIGNED;
            e->positions[2 * idx] = (int16_t)nr; e->positions[2 * idx + 1] = (int16_t)nc;
            e->occ_owner[nr * C + nc] = (int16_t)idx;
        }
        emit_agent_obs(e, idx, out); /* ENV:528-536: staggered snapshot (F3) */

        int on_goal = e->positions[2 * idx] == e->goals[2 * idx] &&
                      e->positions[2 * idx + 1] == e->goals[2 * idx + 1];
        e->reached_goal[idx] = (uint8_t)on_goal;
        if (!on_goal) continue;
        if (e->cfg.lifelong_mapf) { /* ENV:547-556 */
            reward[idx] += 0.5f;
            e->goal_step_flags[idx] = 1.0f;
            e->ep_goals_total += 1.0;
            e->completed_once[idx] = 1;
            e->reached[idx] = 0;
            int rc = assign_new_goal(e, idx, goal_rank ? goal_rank[idx] : -1,
                                     goal_override ? goal_override + 2 * idx : NULL);
            if (rc) return rc;
            e->reached_goal[idx] = 0;
            goal_reassigned = 1;
        } else if (!e->reached[idx]) { /* ENV:557-563 */
            e->reached[idx] = 1;
            e->completed_once[idx] = 1;
            reward[idx] += 0.5f;
            e->goal_step_flags[idx] = 1.0f;
            e->ep_goals_total += 1.0;
        }
    }

    if (goal_reassigned) /* ENV:565-575: rebuild every obs from the final state */
        for (int a = 0; a < N; ++a) emit_agent_obs(e, a, out);

    /* lock metrics, ENV:577-606 */
    int deadlock_step = 0, livelock_step = 0;
    double dl_event = 0.0, ll_event = 0.0;
    for (int a = 0; a < N; ++a) { /* flags are cheap; fill them even when metrics are off */
        e->moved[a] = e->positions[2 * a] != e->prev_pos[2 * a] ||
                      e->positions[2 * a + 1] != e->prev_pos[2 * a + 1];
        e->failed[a] = e->actions_taken[a] != 0 && !e->moved[a];
    }
    if (e->cfg.enable_lock_metrics) {
        for (int a = 0; a < N; ++a) {
            int prev_on = e->prev_pos[2 * a] == e->goals[2 * a] &&
                          e->prev_pos[2 * a + 1] == e->goals[2 * a + 1];
            e->prev_on_goal[a] = e->cfg.lifelong_mapf ? 0 : (uint8_t)prev_on; /* ENV:584-587 */
            e->cur_on_goal[a] = e->positions[2 * a] == e->goals[2 * a] &&
                                e->positions[2 * a + 1] == e->goals[2 * a + 1];
            e->gprog[a] = e->cfg.lifelong_mapf ? (e->goal_step_flags[a] > 0.0f)
                                               : (!e->prev_on_goal[a] && e->cur_on_goal[a]);
            e->dist[a] = (int16_t)(abs(e->goals[2 * a] - e->positions[2 * a]) +
                                   abs(e->goals[2 * a + 1] - e->positions[2 * a + 1]));
        }
        append_lock_history(e);
        detect_lock_step(e, &deadlock_step, &livelock_step);
        if (deadlock_step) livelock_step = 0;
        dl_event = (deadlock_step && !e->dl_prev) ? 1.0 : 0.0;
        ll_event = (livelock_step && !e->ll_prev) ? 1.0 : 0.0;
        e->dl_prev = deadlock_step; e->ll_prev = livelock_step;
        e->ep_dl_steps += deadlock_step; e->ep_ll_steps += livelock_step;
        e->ep_dl_events += dl_event; e->ep_ll_events += ll_event;
    }

    /* intent-based blocking, ENV:608-625 */
    double blocking_sum = 0.0;
    for (int b = 0; b < N; ++b) {
        if (!e->reached[b]) continue;
        int br = e->positions[2 * b], bc = e->positions[2 * b + 1];
        if (br != e->prev_pos[2 * b] || bc != e->prev_pos[2 * b + 1]) continue;
        for (int o = 0; o < N; ++o) {
            if (o == b || e->reached[o]) continue;
            if (e->intended[2 * o] == br && e->intended[2 * o + 1] == bc) {
                e->blocking_flags[b] = 1.0f;
                break;
            }
        }
    }
    for (int a = 0; a < N; ++a) {
        e->bp_prev[a] = e->blocking_flags[a]; /* ENV:624 */
        blocking_sum += e->blocking_flags[a];
    }
    e->ep_blocking += blocking_sum;

    /* info, ENV:627-656 */
    double goals_step = 0.0, reached_sum = 0.0, completed_sum = 0.0;
    for (int a = 0; a < N; ++a) {
        goals_step += e->goal_step_flags[a];
        reached_sum += e->reached[a];
        completed_sum += e->completed_once[a];
    }
    double goals_total = e->cfg.lifelong_mapf ? e->ep_goals_total : reached_sum;

    /* collision penalty, ENV:658-666 */
    for (int i = 0; i < N; ++i)
        for (int j = i + 1; j < N; ++j)
            if (e->positions[2 * i] == e->positions[2 * j] &&
                e->positions[2 * i + 1] == e->positions[2 * j + 1]) {
                reward[i] -= 1.0f; reward[j] -= 1.0f;
            }

    /* termination, ENV:668-690 */
    int all_reached = 1;
    for (int a = 0; a < N; ++a) all_reached &= e->reached_goal[a];
    int terminated = 0, truncated = 0;
    if (!e->cfg.lifelong_mapf && all_reached) {
        for (int a = 0; a < N; ++a) reward[a] += 1.0f;
        terminated = 1; truncated = 0;
    } else if (e->step_count >= e->cfg.steps_per_episode) {
        for (int a = 0; a < N; ++a)
            if (!e->cfg.lifelong_mapf && !e->reached_goal[a]) reward[a] -= 1.0f;
        terminated = 1; truncated = 1; /* F6: both set */
    }

    if (out) {
        if (out->reward) memcpy(out->reward, reward, sizeof(float) * N);
        if (out->terminated) *out->terminated = (uint8_t)terminated;
        if (out->truncated) *out->truncated = (uint8_t)truncated;
        if (out->blocking) memcpy(out->blocking, e->blocking_flags, sizeof(float) * N);
        if (out->goal_reached_step) memcpy(out->goal_reached_step, e->goal_step_flags, sizeof(float) * N);
        if (out->moved) memcpy(out->moved, e->moved, N);
        if (out->failed_move) memcpy(out->failed_move, e->failed, N);
        if (out->intended_next) memcpy(out->intended_next, e->intended, sizeof(int16_t) * 2 * N);
        if (out->goal_reassigned) *out->goal_reassigned = (uint8_t)goal_reassigned;
        if (out->info_all) {
            double *ia = out->info_all;
            ia[ORACLE_INFO_GOALS_REACHED_STEP] = goals_step;
            ia[ORACLE_INFO_GOALS_REACHED_TOTAL] = goals_total;
            ia[ORACLE_INFO_BLOCKING_COUNT_STEP] = blocking_sum;
            ia[ORACLE_INFO_BLOCKING_COUNT_TOTAL] = e->ep_blocking;
            ia[ORACLE_INFO_DEADLOCK_STEP] = deadlock_step;
            ia[ORACLE_INFO_LIVELOCK_STEP] = livelock_step;
            ia[ORACLE_INFO_DEADLOCK_EVENT_STEP] = dl_event;
            ia[ORACLE_INFO_LIVELOCK_EVENT_STEP] = ll_event;
            ia[ORACLE_INFO_DEADLOCK_EVENTS_TOTAL] = e->ep_dl_events;
            ia[ORACLE_INFO_LIVELOCK_EVENTS_TOTAL] = e->ep_ll_events;
            ia[ORACLE_INFO_DEADLOCK_STEPS_TOTAL] = e->ep_dl_steps;
            ia[ORACLE_INFO_LIVELOCK_STEPS_TOTAL] = e->ep_ll_steps;
            ia[ORACLE_INFO_COMPLETION_RATIO] = completed_sum / (double)N; /* ENV:638 */
            ia[ORACLE_INFO_THROUGHPUT] =
                goals_total / (double)(e->step_count > 1 ? e->step_count : 1); /* ENV:655 */
        }
    }
    return ORACLE_OK;
}

/* ---------------------------------------------------------------- read-back */
void oracle_get_state(const oracle_env *e, int16_t *positions, int16_t *starts, int16_t *goals,
                      uint8_t *reached, uint8_t *completed_once, float *blocking_prev,
                      int32_t *step_count, double *episode_counters) {
    int N = e->N;
    if (positions) memcpy(positions, e->positions, sizeof(int16_t) * 2 * N);
    if (starts) memcpy(starts, e->starts, sizeof(int16_t) * 2 * N);
    if (goals) memcpy(goals, e->goals, sizeof(int16_t) * 2 * N);
    if (reached) memcpy(reached, e->reached, N);
    if (completed_once) memcpy(completed_once, e->completed_once, N);
    if (blocking_prev) memcpy(blocking_prev, e->bp_prev, sizeof(float) * N);
    if (step_count) *step_count = e->step_count;
    if (episode_counters) {
        episode_counters[0] = e->ep_goals_total; episode_counters[1] = e->ep_blocking;
        episode_counters[2] = e->ep_dl_events;   episode_counters[3] = e->ep_ll_events;
        episode_counters[4] = e->ep_dl_steps;    episode_counters[5] = e->ep_ll_steps;
    }
}

void oracle_get_owner_grids(const oracle_env *e, int16_t *occ, int16_t *goal) {
    if (occ) memcpy(occ, e->occ_owner, sizeof(int16_t) * e->R * e->C);
    if (goal) memcpy(goal, e->goal_owner, sizeof(int16_t) * e->R * e->C);
}

void oracle_get_last_candidate_counts(const oracle_env *e, int32_t *out) {
    memcpy(out, e->last_cand, sizeof(int32_t) * e->N);
}

void oracle_get_last_ranks(const oracle_env *e, int32_t *out) {
    memcpy(out, e->last_rank, sizeof(int32_t) * e->N);
}

/* ---------------------------------------------------------------- batched stepping (tests) */
static void offset_outputs(const oracle_outputs *o, int b, int N, int V2, oracle_outputs *r) {
    memset(r, 0, sizeof(*r));
    if (!o) return;
    if (o->local_obs) r->local_obs = o->local_obs + (size_t)b * N * V2;
    if (o->action_mask) r->action_mask = o->action_mask + (size_t)b * N * 5;
    if (o->goal_delta) r->goal_delta = o->goal_delta + (size_t)b * N * 2;
    if (o->goal_distance) r->goal_distance = o->goal_distance + (size_t)b * N;
    if (o->blocking_prev) r->blocking_prev = o->blocking_prev + (size_t)b * N;
    if (o->reward) r->reward = o->reward + (size_t)b * N;
    if (o->terminated) r->terminated = o->terminated + b;
    if (o->truncated) r->truncated = o->truncated + b;
    if (o->blocking) r->blocking = o->blocking + (size_t)b * N;
    if (o->goal_reached_step) r->goal_reached_step = o->goal_reached_step + (size_t)b * N;
    if (o->info_all) r->info_all = o->info_all + (size_t)b * ORACLE_INFO_COUNT;
    if (o->moved) r->moved = o->moved + (size_t)b * N;
    if (o->failed_move) r->failed_move = o->failed_move + (size_t)b * N;
    if (o->intended_next) r->intended_next = o->intended_next + (size_t)b * N * 2;
    if (o->goal_reassigned) r->goal_reassigned = o->goal_reassigned + b;
}

int oracle_reset_many(oracle_env **envs, int num_envs, int mode, const int16_t *starts,
                      const int16_t *goals, const uint8_t *mask, oracle_outputs *out) {
    for (int b = 0; b < num_envs; ++b) {
        if (mask && !mask[b]) continue;
        oracle_env *e = envs[b];
        oracle_outputs o;
        offset_outputs(out, b, e->N, e->V * e->V, &o);
        int rc = oracle_reset(e, mode, starts ? starts + (size_t)b * e->N * 2 : NULL,
                              goals ? goals + (size_t)b * e->N * 2 : NULL, out ? &o : NULL);
        if (rc) return rc;
    }
    return ORACLE_OK;
}

int oracle_step_many(oracle_env **envs, int num_envs, const int8_t *actions, const int32_t *goal_rank,
                     const int16_t *goal_override, oracle_outputs *out, int32_t *ranks_out) {
    for (int b = 0; b < num_envs; ++b) {
        oracle_env *e = envs[b];
        int N = e->N;
        oracle_outputs o;
        offset_outputs(out, b, N, e->V * e->V, &o);
        int rc = oracle_step(e, actions ? actions + (size_t)b * N : NULL,
                             goal_rank ? goal_rank + (size_t)b * N : NULL,
                             goal_override ? goal_override + (size_t)b * N * 2 : NULL, out ? &o : NULL);
        if (rc) return rc;
        if (ranks_out) memcpy(ranks_out + (size_t)b * N, e->last_rank, sizeof(int32_t) * N);
    }
    return ORACLE_OK;
}

void oracle_get_state_many(oracle_env **envs, int num_envs, int16_t *positions, int16_t *starts,
                           int16_t *goals, uint8_t *reached, uint8_t *completed_once,
                           float *blocking_prev, int32_t *step_count, double *episode_counters) {
    for (int b = 0; b < num_envs; ++b) {
        int N = envs[b]->N;
        oracle_get_state(envs[b], positions ? positions + (size_t)b * N * 2 : NULL,
                         starts ? starts + (size_t)b * N * 2 : NULL,
                         goals ? goals + (size_t)b * N * 2 : NULL, reached ? reached + (size_t)b * N : NULL,
                         completed_once ? completed_once + (size_t)b * N : NULL,
                         blocking_prev ? blocking_prev + (size_t)b * N : NULL,
                         step_count ? step_count + b : NULL,
                         episode_counters ? episode_counters + (size_t)b * 6 : NULL);
    }
}

void oracle_sample_actions_philox_many(oracle_env **envs, int num_envs, uint64_t call_counter, int masked,
                                       const int8_t *mask, int8_t *actions) {
    for (int b = 0; b < num_envs; ++b) {
        int N = envs[b]->N;
        oracle_sample_actions_philox(envs[b], call_counter, masked, mask ? mask + (size_t)b * N * 5 : NULL,
                                     actions + (size_t)b * N);
    }
}

/* ---------------------------------------------------------------- flat obs, ENV:214-265,306-328 */
int oracle_flat_obs_dim(const oracle_config *cfg, int gdist, int bp, int mask) {
    int V = 2 * cfg->sensor_range + 1;
    return V * V + 2 + (gdist ? 1 : 0) + (bp ? 1 : 0) + (mask ? 5 : 0);
}

void oracle_pack_flat_obs(const oracle_config *cfg, const oracle_outputs *o, int gdist, int bp,
                          int mask, float *flat) {
    int V2 = (2 * cfg->sensor_range + 1) * (2 * cfg->sensor_range + 1);
    int D = oracle_flat_obs_dim(cfg, gdist, bp, mask);
    for (int a = 0; a < cfg->num_agents; ++a) {
        float *f = flat + (size_t)a * D;
        int k = 0;
        for (int i = 0; i < V2; ++i) f[k++] = (float)o->local_obs[(size_t)a * V2 + i];
        f[k++] = o->goal_delta[2 * a];
        f[k++] = o->goal_delta[2 * a + 1];
        if (gdist) f[k++] = o->goal_distance[a];
        if (bp) f[k++] = o->blocking_prev[a];
        if (mask) for (int i = 0; i < 5; ++i) f[k++] = (float)o->action_mask[5 * a + i];
    }
}

/* ---------------------------------------------------------------- threaded benchmark loop */
typedef struct {
    oracle_env **envs;
    int begin, end, steps, mode, deterministic;
    uint64_t seed;
    int64_t env_steps, episodes;
    uint64_t checksum;
} bench_job;

static void *bench_worker(void *arg) {
    bench_job *j = (bench_job *)arg;
    uint64_t rng = j->seed;
    uint64_t cs = 0;
    for (int b = j->begin; b < j->end; ++b) {
        oracle_env *e = j->envs[b];
        int N = e->N, V2 = e->V * e->V;
        uint8_t *obs = (uint8_t *)malloc((size_t)N * V2);
        int8_t *mask = (int8_t *)malloc((size_t)N * 5);
        float *gd = (float *)malloc(sizeof(float) * 2 * N);
        float *bp = (float *)malloc(sizeof(float) * N);
        float *rew = (float *)malloc(sizeof(float) * N);
        int8_t *act = (int8_t *)malloc(N);
        double info_all[ORACLE_INFO_COUNT];
        uint8_t term = 0, trunc = 0;
        oracle_outputs o;
        memset(&o, 0, sizeof(o));
        o.local_obs = obs; o.action_mask = mask; o.goal_delta = gd; o.blocking_prev = bp;
        o.reward = rew; o.terminated = &term; o.truncated = &trunc; o.info_all = info_all;
        oracle_reset(e, j->deterministic ? 0 : 2, NULL, NULL, &o);
        for (int s = 0; s < j->steps; ++s) {
            for (int a = 0; a < N; ++a) {
                if (j->mode == 0) {
                    act[a] = (int8_t)rng_below(&rng, 5);
                } else {
                    int valid[5], nv = 0;
                    for (int k = 0; k < 5; ++k) if (mask[5 * a + k]) valid[nv++] = k;
                    act[a] = nv ? (int8_t)valid[rng_below(&rng, (uint32_t)nv)] : 0;
                }
            }
            oracle_step(e, act, NULL, NULL, &o);
            j->env_steps++;
            for (int k = 0; k < N * V2; ++k) cs = cs * 31u + obs[k];
            for (int a = 0; a < N; ++a) cs += (uint64_t)(int64_t)(rew[a] * 2.0f);
            if (term || trunc) {
                j->episodes++;
                oracle_reset(e, j->deterministic ? 0 : 2, NULL, NULL, &o);
            }
        }
        free(obs); free(mask); free(gd); free(bp); free(rew); free(act);
    }
    j->checksum = cs;
    return NULL;
}

int64_t oracle_bench_run(oracle_env **envs, int num_envs, int steps, int mode, int deterministic,
                         uint64_t action_seed, int threads, int64_t *episodes,
                         uint64_t *checksum) {
    if (threads < 1) threads = 1;
    if (threads > num_envs) threads = num_envs;
    bench_job *jobs = (bench_job *)calloc((size_t)threads, sizeof(bench_job));
    pthread_t *tids = (pthread_t *)calloc((size_t)threads, sizeof(pthread_t));
    for (int t = 0; t < threads; ++t) {
        jobs[t].envs = envs;
        jobs[t].begin = (int)((int64_t)num_envs * t / threads);
        jobs[t].end = (int)((int64_t)num_envs * (t + 1) / threads);
        jobs[t].steps = steps; jobs[t].mode = mode; jobs[t].deterministic = deterministic;
        jobs[t].seed = action_seed * 0x9E3779B97F4A7C15ull + (uint64_t)t * 7919u + 1u;
        pthread_create(&tids[t], NULL, bench_worker, &jobs[t]);
    }
    int64_t total = 0, eps = 0;
    uint64_t cs = 0;
    for (int t = 0; t < threads; ++t) {
        pthread_join(tids[t], NULL);
        total += jobs[t].env_steps; eps += jobs[t].episodes; cs ^= jobs[t].checksum;
    }
    if (episodes) *episodes = eps;
    if (checksum) *checksum = cs;
    free(jobs); free(tids);
    return total;
}
