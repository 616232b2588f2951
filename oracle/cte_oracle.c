/*
 * cte_oracle.c -- CPU oracle (TEST INFRASTRUCTURE ONLY) of the reference's single-agent ("CTE")
 * view of the MAPF grid: src/environments/reference_model_single_agent.py, cited "CTE:a-b".
 * A literal scalar restatement: per-agent position lists, the same loop order and the same
 * order of floating-point additions on the scalar reward (a Python float = C double).
 *
 * Parity status: PINNED against traces recorded from the live Python reference
 * (tests/golden/cte_*.npz, produced by tests/golden/make_golden_cte.py): tests/test_cte_oracle_golden.py.
 */
#include <stdint.h>
#include <string.h>

typedef struct {
    int32_t rows, cols, num_agents, steps_per_episode;
    double blocking_penalty, move_after_goal_penalty; /* CTE:92-93 */
} cte_config;

/* state of ONE env, caller-owned */
typedef struct {
    int16_t *positions; /* [N,2] (row, col) */
    int16_t *goals;     /* [N,2] */
    uint8_t *reached_once; /* [N] goal_reached_once, CTE:91 */
    int32_t *step_count;   /* [1] */
    double *blocking_total; /* [1] _episode_blocking_count, CTE:94 */
} cte_state;

/* CTE:428-441 full-grid observation: obstacles 1, goal of agent i = 2i+3, agent i = 2i+2 (agents overwrite goals) */
void cte_oracle_get_obs(const cte_config *c, const uint8_t *grid, const cte_state *s, uint8_t *obs) {
    memcpy(obs, grid, (size_t)c->rows * c->cols);
    for (int i = 0; i < c->num_agents; ++i) obs[s->goals[2 * i] * c->cols + s->goals[2 * i + 1]] = (uint8_t)(2 * i + 3);
    for (int i = 0; i < c->num_agents; ++i) obs[s->positions[2 * i] * c->cols + s->positions[2 * i + 1]] = (uint8_t)(2 * i + 2);
}

/* CTE:466-489: a move is allowed iff the neighbour is inside the map and its cell value is 0 or odd
 * (odd = goal ... and obstacle, value 1: the reference's rule, kept as is) */
void cte_oracle_get_action_mask(const cte_config *c, const cte_state *s, const uint8_t *obs, int8_t *mask) {
    const int R = c->rows, C = c->cols;
    memset(mask, 0, (size_t)5 * c->num_agents);
    for (int i = 0; i < c->num_agents; ++i) {
        const int x = s->positions[2 * i], y = s->positions[2 * i + 1];
        mask[5 * i] = 1;
        if (x > 0 && (obs[(x - 1) * C + y] == 0 || obs[(x - 1) * C + y] % 2 == 1)) mask[5 * i + 1] = 1;
        if (y < C - 1 && (obs[x * C + y + 1] == 0 || obs[x * C + y + 1] % 2 == 1)) mask[5 * i + 2] = 1;
        if (x < R - 1 && (obs[(x + 1) * C + y] == 0 || obs[(x + 1) * C + y] % 2 == 1)) mask[5 * i + 3] = 1;
        if (y > 0 && (obs[x * C + y - 1] == 0 || obs[x * C + y - 1] % 2 == 1)) mask[5 * i + 4] = 1;
    }
}

/* CTE:218-235 (everything but the layout draw, which the caller does) */
void cte_oracle_reset(const cte_config *c, cte_state *s) {
    *s->step_count = 0;
    *s->blocking_total = 0.0;
    memset(s->reached_once, 0, (size_t)c->num_agents);
}

/* CTE:237-346.  info = [blocking_count_step, goals_reached_step, goals_reached_total, blocking_count_total].
 * Returns 0, or -1 for an invalid action (ValueError at CTE:375-377; state is left as the reference leaves it). */
int cte_oracle_step(const cte_config *c, const uint8_t *grid, cte_state *s, const int8_t *action, uint8_t *obs,
                    int8_t *mask, double *reward_out, uint8_t *terminated, uint8_t *truncated, double *info) {
    const int N = c->num_agents, R = c->rows, C = c->cols;
    int16_t prev[64], intended[64];
    uint8_t reached_goal[32];
    if (N > 32) return -4;
    *s->step_count += 1;
    double reward = 0.0;
    memcpy(prev, s->positions, sizeof(int16_t) * 2 * N);
    double blocking_count_step = 0.0, goals_reached_step = 0.0;
    for (int i = 0; i < N; ++i) {
        reached_goal[i] = 0;
        int ny = s->positions[2 * i], nx = s->positions[2 * i + 1];
        switch (action[i]) { /* CTE:348-379 */
            case 0: break;
            case 1: ny -= 1; break;
            case 2: nx += 1; break;
            case 3: ny += 1; break;
            case 4: nx -= 1; break;
            default: return -1;
        }
        intended[2 * i] = (int16_t)ny; intended[2 * i + 1] = (int16_t)nx;
        int ok = ny >= 0 && ny < R && nx >= 0 && nx < C && grid[ny * C + nx] == 0; /* CTE:256-263 */
        if (ok)
            for (int j = 0; j < N; ++j)
                if (j != i && s->positions[2 * j] == ny && s->positions[2 * j + 1] == nx) { ok = 0; break; }
        if (ok) { s->positions[2 * i] = (int16_t)ny; s->positions[2 * i + 1] = (int16_t)nx; }
        if (s->positions[2 * i] == s->goals[2 * i] && s->positions[2 * i + 1] == s->goals[2 * i + 1]) { /* CTE:271-276 */
            reached_goal[i] = 1;
            if (!s->reached_once[i]) { s->reached_once[i] = 1; reward += 0.5; goals_reached_step += 1.0; }
        }
    }
    cte_oracle_get_obs(c, grid, s, obs);
    cte_oracle_get_action_mask(c, s, obs, mask);
    for (int i = 0; i < N; ++i) /* CTE:285-290 */
        for (int j = i + 1; j < N; ++j)
            if (s->positions[2 * i] == s->positions[2 * j] && s->positions[2 * i + 1] == s->positions[2 * j + 1]) reward -= 1;
    for (int b = 0; b < N; ++b) { /* CTE:292-307 intent-based local blocking penalty */
        if (!s->reached_once[b]) continue;
        if (s->positions[2 * b] != prev[2 * b] || s->positions[2 * b + 1] != prev[2 * b + 1]) continue;
        for (int o = 0; o < N; ++o) {
            if (o == b || s->reached_once[o]) continue;
            if (intended[2 * o] == s->positions[2 * b] && intended[2 * o + 1] == s->positions[2 * b + 1]) {
                reward += c->blocking_penalty;
                blocking_count_step += 1.0;
                break;
            }
        }
    }
    *s->blocking_total += blocking_count_step;
    for (int i = 0; i < N; ++i) { /* CTE:309-315 */
        if (!s->reached_once[i]) continue;
        if (s->positions[2 * i] != prev[2 * i] || s->positions[2 * i + 1] != prev[2 * i + 1]) reward += c->move_after_goal_penalty;
    }
    int all = 1;
    for (int i = 0; i < N; ++i) all &= reached_goal[i];
    if (all) { reward += N; *terminated = 1; *truncated = 0; } /* CTE:317-320 */
    else if (*s->step_count >= c->steps_per_episode) {         /* CTE:329-337 */
        for (int i = 0; i < N; ++i) if (!reached_goal[i]) reward -= 1;
        *terminated = 1; *truncated = 1;
    } else { *terminated = 0; *truncated = 0; }
    double once = 0.0;
    for (int i = 0; i < N; ++i) once += s->reached_once[i] ? 1.0 : 0.0;
    info[0] = blocking_count_step; info[1] = goals_reached_step; info[2] = once; info[3] = *s->blocking_total;
    *reward_out = reward;
    return 0;
}
