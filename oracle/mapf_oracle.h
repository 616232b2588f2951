/*
 * mapf_oracle.h -- CPU oracle for the MAPF environment transition.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the reference's
 * reset()/step() path (src/environments/reference_model_multi_agent.py, "ENV" below)
 * kept deliberately literal: owner grids, ring buffers and the sequential agent loop
 * are the reference's own data structures, not the GPU formulation.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it; the product package never does.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks it against
 *   (1) the reference's two SHA-256 golden trace digests
 *       (tests/test_reference_model_multi_agent_parity.py:12,19),
 *   (2) the reference's known-answer tests (blocking delay, deadlock edge, lifelong),
 *   (3) traces recorded from the live Python reference in the build container
 *       (npz files under tests/golden/, produced by tests/golden/make_golden.py).
 */
#ifndef MAPF_ORACLE_H
#define MAPF_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ENV:38-61 -- the env_config keys that affect the transition. */
typedef struct {
    int32_t rows, cols;
    int32_t num_agents;
    int32_t sensor_range;
    int32_t steps_per_episode;
    int32_t lifelong_mapf;
    int32_t enable_lock_metrics;
    int32_t deadlock_window_steps;
    int32_t livelock_window_steps;
    int32_t lock_nearby_manhattan;
    int32_t lock_min_neighbors;
    double  lock_progress_epsilon;
    int32_t normalize_goal_delta;
} oracle_config;

/* info["__all__"] slots (ENV:639-655), in this order. */
enum {
    ORACLE_INFO_GOALS_REACHED_STEP = 0,
    ORACLE_INFO_GOALS_REACHED_TOTAL,
    ORACLE_INFO_BLOCKING_COUNT_STEP,
    ORACLE_INFO_BLOCKING_COUNT_TOTAL,
    ORACLE_INFO_DEADLOCK_STEP,
    ORACLE_INFO_LIVELOCK_STEP,
    ORACLE_INFO_DEADLOCK_EVENT_STEP,
    ORACLE_INFO_LIVELOCK_EVENT_STEP,
    ORACLE_INFO_DEADLOCK_EVENTS_TOTAL,
    ORACLE_INFO_LIVELOCK_EVENTS_TOTAL,
    ORACLE_INFO_DEADLOCK_STEPS_TOTAL,
    ORACLE_INFO_LIVELOCK_STEPS_TOTAL,
    ORACLE_INFO_COMPLETION_RATIO, /* lifelong only in the reference; always filled here */
    ORACLE_INFO_THROUGHPUT,       /* lifelong only in the reference; always filled here */
    ORACLE_INFO_COUNT
};

/* Per-step / per-reset outputs for ONE env.  Any pointer may be NULL. */
typedef struct {
    uint8_t *local_obs;         /* [N, V, V]  ENV:707-747 */
    int8_t  *action_mask;       /* [N, 5]     ENV:749-773 */
    float   *goal_delta;        /* [N, 2]     ENV:330-335 */
    float   *goal_distance;     /* [N]        ENV:320 */
    float   *blocking_prev;     /* [N]        value packed into the obs, ENV:322 */
    float   *reward;            /* [N] */
    uint8_t *terminated;        /* [1] (per-agent values equal "__all__", ENV:668-690) */
    uint8_t *truncated;         /* [1] */
    float   *blocking;          /* [N]  info[aid]["blocking"] */
    float   *goal_reached_step; /* [N]  info[aid]["goal_reached_step"] */
    double  *info_all;          /* [ORACLE_INFO_COUNT] */
    uint8_t *moved;             /* [N]  ENV:582 (filled even when lock metrics are off) */
    uint8_t *failed_move;       /* [N]  ENV:583 */
    int16_t *intended_next;     /* [N, 2] ENV:514-515 */
    uint8_t *goal_reassigned;   /* [1]  ENV:556 */
} oracle_outputs;

typedef struct oracle_env oracle_env;

/* error codes */
#define ORACLE_OK 0
#define ORACLE_ERR_INVALID_ACTION (-1) /* ENV:504-506 ValueError */
#define ORACLE_ERR_NO_GOAL_CELL   (-2) /* ENV:296-298 RuntimeError */
#define ORACLE_ERR_TOO_FEW_CELLS  (-3) /* ENV:270-275 ValueError */
#define ORACLE_ERR_BAD_ARG        (-4)

oracle_env *oracle_create(const oracle_config *cfg, const uint8_t *grid, uint64_t seed);
void        oracle_destroy(oracle_env *env);

int oracle_num_free(const oracle_env *env);
/* free cells in np.argwhere (row-major) order, ENV:82 */
void oracle_get_free_positions(const oracle_env *env, int16_t *out /* [F,2] */);

/* ENV:124-132 / ENV:278-282: install starts+goals, positions <- starts, rebuild owner grids. */
int oracle_set_layout(oracle_env *env, const int16_t *starts, const int16_t *goals);

/* Private-state injection used by the reference's tests (_set_state helpers):
 * every pointer may be NULL (left untouched).  Owner grids are rebuilt (ENV:200-212). */
int oracle_set_state(oracle_env *env, const int16_t *positions, const int16_t *starts,
                     const int16_t *goals, const uint8_t *reached, const uint8_t *completed_once,
                     const float *blocking_prev, const int32_t *step_count,
                     const double *episode_goals_total);
void oracle_reset_lock_tracking(oracle_env *env); /* ENV:360-372 */

/* ENV:440-472.  mode 0: deterministic (positions <- starts, goals kept, F7)
 *               mode 1: layout given by starts/goals (what rng.choice produced)
 *               mode 2: layout drawn from the oracle's own RNG (uniform w/o replacement) */
int oracle_reset(oracle_env *env, int mode, const int16_t *starts, const int16_t *goals,
                 oracle_outputs *out);

/* ENV:474-695.  goal_rank[N]: if >=0 it replaces rng.integers(n) at ENV:300 for that agent;
 * goal_override[N,2]: if row>=0 it replaces the whole selection.  Both may be NULL. */
int oracle_step(oracle_env *env, const int8_t *actions, const int32_t *goal_rank,
                const int16_t *goal_override, oracle_outputs *out);

/* state read-back */
void oracle_get_state(const oracle_env *env, int16_t *positions, int16_t *starts, int16_t *goals,
                      uint8_t *reached, uint8_t *completed_once, float *blocking_prev,
                      int32_t *step_count, double *episode_counters /* [6]: goals, blocking,
                      dl_events, ll_events, dl_steps, ll_steps */);
void oracle_get_owner_grids(const oracle_env *env, int16_t *occupancy_owner, int16_t *goal_owner);
/* number of candidate cells the last reassignment of each agent saw (debug), [N] */
void oracle_get_last_candidate_counts(const oracle_env *env, int32_t *out);

/* rank (index into the candidate list, ENV:300) used by the last step's reassignment of each
 * agent, -1 where none happened; lets a recorded oracle run be replayed through goal_rank. [N] */
void oracle_get_last_ranks(const oracle_env *env, int32_t *out);

/* Device-RNG replay (no reference counterpart: the reference draws with numpy PCG64, the kernels with
 * Philox4x32-10 keyed by (seed, global env id); see include/mapf_b200.h).  With `on` != 0 every later
 * layout draw (oracle_reset mode 2) and lifelong goal draw (oracle_step without rank/override) follows
 * the kernels' documented draw discipline, starting at the env's RNG counter `counter`. */
void oracle_set_philox(oracle_env *env, int on, uint64_t seed, int64_t env_global, uint32_t counter);
uint32_t oracle_get_philox_counter(const oracle_env *env);
void oracle_sample_actions_philox(const oracle_env *env, uint64_t call_counter, int masked,
                                  const int8_t *mask /* [N,5] */, int8_t *actions /* [N] */);
void oracle_sample_actions_philox_many(oracle_env **envs, int num_envs, uint64_t call_counter, int masked,
                                       const int8_t *mask /* [B,N,5] */, int8_t *actions /* [B,N] */);
/* one Philox4x32-10 block (known-answer tests) */
void oracle_philox_raw(uint32_t k0, uint32_t k1, uint32_t ctr[4]);

/* Batched variants for parity tests at B > 1: env b uses slice b of every [B, ...] array. */
int oracle_reset_many(oracle_env **envs, int num_envs, int mode, const int16_t *starts,
                      const int16_t *goals, const uint8_t *mask, oracle_outputs *out);
int oracle_step_many(oracle_env **envs, int num_envs, const int8_t *actions, const int32_t *goal_rank,
                     const int16_t *goal_override, oracle_outputs *out, int32_t *ranks_out);
void oracle_get_state_many(oracle_env **envs, int num_envs, int16_t *positions, int16_t *starts,
                           int16_t *goals, uint8_t *reached, uint8_t *completed_once,
                           float *blocking_prev, int32_t *step_count, double *episode_counters);

/* ENV:306-328: pack channels into the flat float32 vector.  Returns D. */
int oracle_flat_obs_dim(const oracle_config *cfg, int include_goal_distance,
                        int include_blocking_pressure, int include_action_mask);
void oracle_pack_flat_obs(const oracle_config *cfg, const oracle_outputs *o,
                          int include_goal_distance, int include_blocking_pressure,
                          int include_action_mask, float *flat /* [N, D] */);

/* ------------------------------------------------------------------------------------------
 * Batched helpers (bench.py cpu_baseline / --impl reference leg; also parity tests at B>1).
 * envs[b] are independent; work is split over `threads` POSIX threads.
 * mode: 0 = uniform random actions (scripts/benchmark_multi_agent_env.py:38-39),
 *       1 = uniform over the valid action mask (scripts/benchmark_multi_agent_env.py:42-57).
 * Every env is stepped `steps` times; a finished episode is followed by reset(mode 2 or 0)
 * inside the loop exactly like run_benchmark (scripts/benchmark_multi_agent_env.py:89-95).
 * Returns total env-steps executed; *episodes gets the number of finished episodes and
 * *checksum a value depending on every output byte (keeps the compiler honest). */
int64_t oracle_bench_run(oracle_env **envs, int num_envs, int steps, int mode, int deterministic,
                         uint64_t action_seed, int threads, int64_t *episodes,
                         uint64_t *checksum);

#ifdef __cplusplus
}
#endif
#endif
