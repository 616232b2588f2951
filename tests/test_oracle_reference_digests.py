"""The oracle reproduces the reference's own golden vectors and known-answer tests
(SURVEY §8c items 1-5), without the reference being present."""
import numpy as np
import pytest

import refcompat
from dl_reference_models_b200 import maps
from dl_reference_models_b200.actions import LEFT, NO_OP, RIGHT
from oracle.oracle import OracleEnv, flat_obs


def test_stochastic_sha256_digest():
    # reference: tests/test_reference_model_multi_agent_parity.py:12-17,141-144
    env = refcompat.OracleDictEnv(refcompat.golden_env_config(False))
    digest, summary = refcompat.trace_digest(env)
    assert summary == refcompat.EXPECTED_STOCHASTIC_SUMMARY
    assert digest == refcompat.EXPECTED_STOCHASTIC_DIGEST


def test_deterministic_sha256_digest():
    # reference: tests/test_reference_model_multi_agent_parity.py:19-24,147-150
    env = refcompat.OracleDictEnv(refcompat.golden_env_config(True))
    digest, summary = refcompat.trace_digest(env)
    assert summary == refcompat.EXPECTED_DETERMINISTIC_SUMMARY
    assert digest == refcompat.EXPECTED_DETERMINISTIC_DIGEST


def _lock_env(**kw):
    cfg = {"num_agents": 2, "steps_per_episode": 50, "sensor_range": 1, "deadlock_window_steps": 2,
           "livelock_window_steps": 4, "lock_nearby_manhattan": 2, "lock_progress_epsilon": 1,
           "lock_min_neighbors": 1}
    cfg.update(kw)
    env = OracleEnv(cfg, maps.get_grid("ReferenceModel-1-3"))
    env.reset(mode=2)
    return env


def _inject(env, positions, goals):
    n = len(positions)
    env.set_state(positions=positions, starts=positions, goals=goals, reached=[0] * n,
                  completed_once=[0] * n, blocking_prev=[0.0] * n, step_count=0,
                  episode_goals_total=0.0)
    env.reset_lock_tracking()


def test_deadlock_fires_at_step_two_for_on_goal_blocker():
    # reference: tests/test_reference_model_lock_metrics.py:44-63
    env = _lock_env()
    _inject(env, [(2, 0), (2, 1)], [(2, 2), (2, 1)])
    infos = [env.step([RIGHT, NO_OP]).info_dict(False) for _ in range(3)]
    assert infos[0]["deadlock_event_step"] == 0.0
    assert infos[1]["deadlock_step"] == 1.0 and infos[1]["deadlock_event_step"] == 1.0
    assert infos[1]["livelock_step"] == 0.0 and infos[1]["deadlock_events_total"] == 1.0
    assert infos[2]["deadlock_event_step"] == 0.0


def test_deadlock_uses_current_state_not_sticky_flags():
    # reference: tests/test_reference_model_lock_metrics.py:66-86
    env = _lock_env()
    _inject(env, [(2, 0), (2, 2)], [(2, 1), (4, 2)])
    env.step([RIGHT, NO_OP])
    assert env.state()["completed_once"][0] == 1
    env.step([LEFT, LEFT])
    assert tuple(env.state()["positions"][0]) == (2, 0)
    env.step([RIGHT, NO_OP])
    info = env.step([RIGHT, NO_OP]).info_dict(False)
    assert info["deadlock_step"] == 1.0 and info["deadlock_event_step"] == 1.0


def test_blocking_pressure_is_delayed_by_one_step():
    # reference: tests/test_reference_model_multi_agent_invariants.py:119-147
    env = _lock_env(steps_per_episode=20, deadlock_window_steps=8, livelock_window_steps=16)
    _inject(env, [(2, 0), (2, 1)], [(2, 2), (2, 1)])
    seq = [env.step(a).blocking_prev[1] for a in ([RIGHT, NO_OP], [RIGHT, NO_OP], [NO_OP, NO_OP],
                                                  [NO_OP, NO_OP])]
    assert seq == [0.0, 1.0, 1.0, 0.0]


def test_lifelong_reassignment_and_ratios():
    # reference: tests/test_reference_model_lifelong.py:73-94,176-194
    cfg = {"num_agents": 2, "steps_per_episode": 20, "sensor_range": 2, "lifelong_mapf": True}
    env = OracleEnv(cfg, maps.get_grid("ReferenceModel-2-1"), seed=3)
    env.reset(mode=2)
    _inject(env, [(0, 0), (3, 0)], [(0, 1), (3, 1)])
    r = env.step([RIGHT, NO_OP])
    st = env.state()
    assert r.goal_reached_step[0] == 1.0 and r.reward[0] == 0.5
    new_goal = tuple(st["goals"][0])
    assert new_goal != (0, 1) and new_goal != tuple(st["goals"][1])
    assert new_goal not in {tuple(p) for p in st["positions"]}
    info = r.info_dict(True)
    assert info["completion_ratio"] == pytest.approx(0.5)
    assert info["throughput"] == pytest.approx(info["goals_reached_total"] / st["step_count"][0])
    assert r.terminated[0] == 0 and r.truncated[0] == 0
    assert env.last_candidate_counts()[0] == 116 - 2 - 1  # free - occupied - other goal


def test_obs_priority_known_answer():
    # reference: tests/get_obs.py:145-167 (same priorities: obstacle>agent>own goal>other goal)
    grid = np.array([[1, 0, 0], [0, 0, 0], [1, 0, 1]], np.uint8)
    env = OracleEnv({"num_agents": 2, "sensor_range": 1}, grid)
    env.set_layout([(1, 1), (1, 0)], [(1, 2), (0, 2)])
    r = env.reset(mode=0)
    assert r.local_obs[0].tolist() == [[1, 0, 4], [2, 0, 3], [1, 0, 1]]
    assert r.action_mask[0].tolist() == [1, 1, 1, 1, 0]
    assert r.local_obs[1].tolist() == [[1, 1, 0], [1, 0, 2], [1, 1, 0]]
    assert r.action_mask[1].tolist() == [1, 0, 0, 0, 0]


@pytest.mark.parametrize("flags", [(a, b, c) for a in (0, 1) for b in (0, 1) for c in (0, 1)])
@pytest.mark.parametrize("normalize", [True, False])
def test_flat_obs_layout_and_dtype(flags, normalize):
    # reference: tests/test_reference_model_observation_dtypes.py:58-101 and ENV:214-265
    cfg = {"num_agents": 4, "sensor_range": 2, "steps_per_episode": 20, "normalize_goal_delta": normalize}
    env = OracleEnv(cfg, maps.get_grid("ReferenceModel-2-1"))
    s = maps.get_start_positions("ReferenceModel-2-1", 4)
    g = maps.get_goal_positions("ReferenceModel-2-1", 4)
    env.set_layout(list(s.values()), list(g.values()))
    for r in (env.reset(mode=0), env.step([0, 0, 0, 0])):
        f = flat_obs(r, *flags)
        assert f.dtype == np.float32 and f.shape == (4, 25 + 2 + sum(flags[:2]) + 5 * flags[2])
        hi = 1.0 if normalize else 19.0
        assert np.all(np.abs(f[:, 25:27]) <= hi) and np.all(f[:, :25] <= 4)
        if flags[2]:
            assert set(np.unique(f[:, -5:])) <= {0.0, 1.0}


def test_error_codes():
    from oracle.oracle import ERR_INVALID_ACTION, ERR_TOO_FEW_CELLS, OracleError

    env = OracleEnv({"num_agents": 2}, maps.get_grid("ReferenceModel-1-3"))
    env.reset(mode=2)
    with pytest.raises(OracleError) as ei:
        env.step([5, 0])  # ENV:504-506
    assert ei.value.code == ERR_INVALID_ACTION
    env = OracleEnv({"num_agents": 6}, maps.get_grid("ReferenceModel-1-1"))  # 10 free < 12
    with pytest.raises(OracleError) as ei:
        env.reset(mode=2)  # ENV:270-275
    assert ei.value.code == ERR_TOO_FEW_CELLS


def test_own_rng_layouts_are_uniform_and_distinct():
    env = OracleEnv({"num_agents": 4, "sensor_range": 1}, maps.get_grid("ReferenceModel-1-4"), seed=9)
    free = {tuple(p) for p in env.free_positions()}
    counts = {}
    for _ in range(2600):
        env.reset(mode=2)
        st = env.state()
        cells = [tuple(p) for p in st["starts"]] + [tuple(p) for p in st["goals"]]
        assert len(set(cells)) == 8 and set(cells) <= free
        counts[cells[0]] = counts.get(cells[0], 0) + 1
    assert len(counts) == 13 and min(counts.values()) > 130  # expectation 200 per cell
