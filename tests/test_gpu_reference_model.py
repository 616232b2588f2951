"""The drop-in ``ReferenceModel`` (RLlib dict API over the CUDA kernels) against the reference's
own golden vectors and behavioural tests, restated (the reference's test files are not present on
the GPU box):
  tests/test_reference_model_multi_agent_parity.py      -> SHA-256 digests, default obs contract
  tests/test_reference_model_multi_agent_invariants.py  -> invariants, blocking delay, lite == full
  tests/test_reference_model_lifelong.py                -> reassignment, monkeypatched goal hook, ratios
  tests/test_reference_model_lock_metrics.py            -> deadlock edge / non-sticky flags
  tests/test_reference_model_observation_dtypes.py      -> 16-way float32 / bounds matrix
and, beyond them, deep dict equality against the oracle-backed dict env incl. all lock keys."""
import numpy as np
import pytest

import refcompat
from dl_reference_models_b200.actions import DOWN, LEFT, NO_OP, RIGHT, UP

pytestmark = pytest.mark.gpu


def RM(cfg):
    from dl_reference_models_b200.reference_model import ReferenceModel

    return ReferenceModel(cfg)


def base_cfg(**kw):
    cfg = {"env_name": "ReferenceModel-2-1", "seed": 123, "deterministic": False, "num_agents": 4,
           "steps_per_episode": 100, "sensor_range": 2, "info_mode": "lite", "training_execution_mode": "CTDE",
           "render_env": False}
    cfg.update(kw)
    return cfg


def set_state(env, positions, goals):
    """What the reference's tests do to its private arrays (invariants.py:28-38, lifelong.py:29-39)."""
    np.copyto(env._positions_arr, np.asarray(positions, dtype=env._coord_dtype))
    np.copyto(env._starts_arr, np.asarray(positions, dtype=env._coord_dtype))
    np.copyto(env._goals_arr, np.asarray(goals, dtype=env._coord_dtype))
    env._rebuild_goal_owner()
    env._rebuild_occupancy_owner()
    env._reached_arr[:] = False
    env._completed_once_arr[:] = False
    env.goal_reached_once = dict.fromkeys(env.agents, False)
    env._blocking_pressure_prev_arr.fill(0.0)
    env._episode_goals_reached_total = 0.0
    env.step_count = 0


# ------------------------------------------------------------------ golden digests
def test_stochastic_sha256_digest():
    digest, summary = refcompat.trace_digest(RM(refcompat.golden_env_config(False)))
    assert summary == refcompat.EXPECTED_STOCHASTIC_SUMMARY
    assert digest == refcompat.EXPECTED_STOCHASTIC_DIGEST


def test_deterministic_sha256_digest():
    digest, summary = refcompat.trace_digest(RM(refcompat.golden_env_config(True)))
    assert summary == refcompat.EXPECTED_DETERMINISTIC_SUMMARY
    assert digest == refcompat.EXPECTED_DETERMINISTIC_DIGEST


def test_default_observation_contract():
    env = RM({"env_name": "ReferenceModel-2-1", "num_agents": 4, "sensor_range": 2, "seed": 1})
    obs, infos = env.reset()
    assert list(env._obs_slices) == ["local_obs", "goal_delta", "blocking_pressure_prev"]
    assert env.observation_space.shape == (28,) and env.observation_space.dtype == np.float32
    assert int(env.action_space.n) == 5 and env._action_mask_space.shape == (5,)
    for aid in env.agents:
        assert obs[aid].dtype == np.float32 and obs[aid].shape == (28,) and infos[aid] == {}
    assert env.get_agent_ids() == set(env.agents)


# ------------------------------------------------------------------ deep equality with the oracle dict env
def _deep_equal(a, b, path=""):
    if isinstance(a, dict):
        assert isinstance(b, dict) and set(a) == set(b), f"{path}: keys {sorted(a, key=str)} vs {sorted(b, key=str)}"
        for k in a:
            _deep_equal(a[k], b[k], f"{path}/{k}")
    elif isinstance(a, np.ndarray):
        assert isinstance(b, np.ndarray) and a.dtype == b.dtype and a.shape == b.shape, f"{path}: array meta"
        assert np.array_equal(a, b), f"{path}: {a} vs {b}"
    else:
        assert type(a) is type(b) and a == b, f"{path}: {a!r} vs {b!r}"


@pytest.mark.parametrize("cfg", [
    base_cfg(info_mode="full", include_action_mask_in_obs=True),
    base_cfg(deterministic=True, info_mode="full", include_goal_distance=True),
    base_cfg(env_name="ReferenceModel-1-4", deadlock_window_steps=2, livelock_window_steps=4, steps_per_episode=60),
    base_cfg(env_name="ReferenceModel-1-2", num_agents=2, sensor_range=1, normalize_goal_delta=False,
             deterministic=True, steps_per_episode=30),
    base_cfg(env_name="ReferenceModel-3-1", num_agents=4, sensor_range=3, deterministic=True, info_mode="full"),
], ids=["full-mask", "det-full-gdist", "cross-locks", "n2-sr1-raw", "sr3"])
def test_dict_payloads_equal_oracle_dict_env(cfg):
    """Non-lifelong configs share the numpy layout stream, so everything must be identical:
    obs, rewards, terminated, truncated and info (including the lock keys the digests drop)."""
    ours, ref = RM(cfg), refcompat.OracleDictEnv(cfg)
    rng = np.random.default_rng(5)
    for ep in range(3):
        (o1, i1), (o2, i2) = ours.reset(), ref.reset()
        _deep_equal(o1, o2, f"ep{ep}/reset/obs")
        _deep_equal(i1, i2, f"ep{ep}/reset/info")
        for s in range(int(cfg["steps_per_episode"]) + 5):
            acts = {a: int(rng.integers(0, 5)) for a in ours.agents}
            r1, r2 = ours.step(acts), ref.step(acts)
            for name, x, y in zip(("obs", "rew", "term", "trunc", "info"), r1, r2):
                _deep_equal(x, y, f"ep{ep}/step{s}/{name}")
            if r1[2]["__all__"] or r1[3]["__all__"]:
                break


# ------------------------------------------------------------------ invariants
def test_unique_starts_goals_and_disjoint_sets():
    env = RM(base_cfg())
    for _ in range(100):
        env.reset()
        starts = [tuple(map(int, env.starts[a])) for a in env.agents]
        goals = [tuple(map(int, env.goals[a])) for a in env.agents]
        assert len(set(starts)) == 4 and len(set(goals)) == 4 and set(starts).isdisjoint(goals)


@pytest.mark.parametrize("backend", ["numpy", "philox"])
def test_positions_stay_in_bounds_and_collision_free(backend):
    env = RM(base_cfg(rng_backend=backend))
    rng = np.random.default_rng(77)
    obs, _ = env.reset()
    for _ in range(600):
        obs, _r, term, trunc, _i = env.step({a: int(rng.integers(0, 5)) for a in env.agents})
        cells = [tuple(map(int, env.positions[a])) for a in env.agents]
        for y, x in cells:
            assert 0 <= y < env.grid.shape[0] and 0 <= x < env.grid.shape[1] and env.grid[y, x] == env.EMPTY_CELL
        assert len(set(cells)) == len(cells)
        if term["__all__"] or trunc["__all__"]:
            obs, _ = env.reset()
    assert obs


def test_action_mask_matches_local_observation():
    env = RM(base_cfg(deterministic=True))
    env.reset()
    ok, c = set(env.TRAVERSABLE_LOCAL_VALUES), env.sensor_range
    for aid in env.agents:
        lo = env.get_obs(aid)
        m = env.get_action_mask(lo)
        assert lo.dtype == np.uint8 and lo.shape == (5, 5) and m.dtype == np.int8
        assert [int(v) for v in m] == [1, int(lo[c - 1, c] in ok), int(lo[c, c + 1] in ok),
                                      int(lo[c + 1, c] in ok), int(lo[c, c - 1] in ok)]


def test_mask_slice_absent_and_pressure_zero_on_reset():
    env = RM(base_cfg(deterministic=True, include_action_mask_in_obs=False))
    assert "action_mask" not in env._obs_slices and "blocking_pressure_prev" in env._obs_slices
    obs, _ = env.reset()
    sl = env._obs_slices["blocking_pressure_prev"]
    for aid in env.agents:
        np.testing.assert_array_equal(obs[aid][sl], np.array([0.0], np.float32))


def test_blocking_pressure_prev_transitions_for_goal_blocker():
    env = RM(base_cfg(env_name="ReferenceModel-1-3", num_agents=2, steps_per_episode=20, sensor_range=1))
    env.reset()
    set_state(env, [(2, 0), (2, 1)], [(2, 2), (2, 1)])
    sl = env._obs_slices["blocking_pressure_prev"]
    block, idle = {"agent_0": RIGHT, "agent_1": NO_OP}, {"agent_0": NO_OP, "agent_1": NO_OP}
    seq = [float(env.step(a)[0]["agent_1"][sl][0]) for a in (block, block, idle, idle)]
    assert seq == [0.0, 1.0, 1.0, 0.0]


def test_info_mode_lite_and_full_payloads_have_same_metrics():
    lite, full = RM(base_cfg(deterministic=True)), RM(base_cfg(deterministic=True, info_mode="full"))
    rng = np.random.default_rng(2026)
    (lo, li), (fo, fi) = lite.reset(), full.reset()
    for a in lite.agents:
        np.testing.assert_array_equal(lo[a], fo[a])
        assert li[a] == {} and {"local_obs", "action_mask", "position", "goal", "goal_delta"} <= set(fi[a])
    for _ in range(120):
        acts = {a: int(rng.integers(0, 5)) for a in lite.agents}
        lo, lr, lt, lu, li = lite.step(acts)
        fo, fr, ft, fu, fi = full.step(acts)
        assert lr == fr and lt == ft and lu == fu and li["__all__"] == fi["__all__"]
        for a in lite.agents:
            np.testing.assert_array_equal(lo[a], fo[a])
            for k in ("blocking", "goal_reached_step", "goals_reached_total", "blocking_count_total"):
                assert li[a][k] == fi[a][k]
            assert "local_obs" not in li[a] and "local_obs" in fi[a]
        if lt["__all__"] or lu["__all__"]:
            lite.reset()
            full.reset()


def test_errors_match_the_reference():
    with pytest.raises(ValueError, match="Unsupported info_mode"):
        RM(base_cfg(info_mode="invalid"))
    with pytest.raises(ValueError, match="Unknown environment name"):
        RM(base_cfg(env_name="nope"))
    with pytest.raises(ValueError, match="free cells"):
        RM(base_cfg(env_name="ReferenceModel-1-1", num_agents=6))  # 10 free < 12
    env = RM(base_cfg())
    env.reset()
    with pytest.raises(ValueError, match="Invalid action 5 for agent_1"):
        env.step({"agent_0": 0, "agent_1": 5, "agent_2": 0, "agent_3": 0})
    before = {a: tuple(map(int, env.positions[a])) for a in env.agents}
    obs, rew, term, trunc, info = env.step({})  # missing actions -> all NO_OP (ENV:498-500)
    assert before == {a: tuple(map(int, env.positions[a])) for a in env.agents}
    with pytest.raises(ValueError, match="Invalid action"):
        env.get_next_position(7, (0, 0))
    assert env.get_next_position(RIGHT, (3, 4)).tolist() == [3, 5]


# ------------------------------------------------------------------ lifelong
def lifelong_cfg(**kw):
    return base_cfg(lifelong_mapf=True, **kw)


def adjacent_pair(env, forbidden):
    free = {tuple(map(int, p)) for p in env._free_positions}
    for src in sorted(free):
        if src in forbidden:
            continue
        for dy, dx in ((-1, 0), (0, 1), (1, 0), (0, -1)):
            dst = (src[0] + dy, src[1] + dx)
            if dst in free and dst not in forbidden:
                return src, dst
    raise RuntimeError("no adjacent free pair")


def towards(src, dst):
    return {(-1, 0): UP, (0, 1): RIGHT, (1, 0): DOWN, (0, -1): LEFT}[(dst[0] - src[0], dst[1] - src[1])]


@pytest.mark.parametrize("backend", ["numpy", "philox"])
def test_reassigns_goal_immediately_after_reach(backend):
    env = RM(lifelong_cfg(num_agents=2, rng_backend=backend))
    env.reset()
    p0 = adjacent_pair(env, set())
    p1 = adjacent_pair(env, {p0[0], p0[1]})
    set_state(env, [p0[0], p1[0]], [p0[1], p1[1]])
    _o, rew, _t, _u, info = env.step({"agent_0": towards(*p0), "agent_1": NO_OP})
    new_goal = tuple(map(int, env.goals["agent_0"]))
    cells = {tuple(map(int, env.positions[a])) for a in env.agents}
    assert info["agent_0"]["goal_reached_step"] == 1.0 and rew["agent_0"] == 0.5
    assert new_goal != p0[1] and new_goal not in cells and new_goal != tuple(map(int, env.goals["agent_1"]))
    ia = info["__all__"]
    assert ia["completion_ratio"] == pytest.approx(0.5)
    assert ia["throughput"] == pytest.approx(ia["goals_reached_total"] / float(env.step_count))


@pytest.mark.parametrize("backend", ["numpy", "philox"])
def test_goals_remain_unique_and_unoccupied(backend):
    env = RM(lifelong_cfg(rng_backend=backend))
    rng = np.random.default_rng(2026)
    env.reset()
    reached = 0.0
    for _ in range(240):
        _o, _r, term, trunc, info = env.step({a: int(rng.integers(0, 5)) for a in env.agents})
        reached += info["__all__"]["goals_reached_step"]
        goals = [tuple(map(int, env.goals[a])) for a in env.agents]
        assert len(set(goals)) == len(goals)
        for (gy, gx), a in zip(goals, env.agents):
            assert env.grid[gy, gx] == env.EMPTY_CELL and (gy, gx) != tuple(map(int, env.positions[a]))
        if term["__all__"] or trunc["__all__"]:
            env.reset()
    assert reached > 0


def test_no_early_termination_in_lifelong():
    env = RM(lifelong_cfg(num_agents=1, steps_per_episode=10))
    env.reset()
    src, dst = adjacent_pair(env, set())
    set_state(env, [src], [dst])
    _o, _r, term, trunc, _i = env.step({"agent_0": towards(src, dst)})
    assert term["__all__"] is False and trunc["__all__"] is False


def test_cumulative_goals_can_exceed_num_agents_with_patched_goal_hook(monkeypatch):
    """The reference's test monkeypatches env._assign_new_goal and reads the private owner grids
    mid-step (tests/test_reference_model_lifelong.py:132-173); the hook must be honoured."""
    env = RM(lifelong_cfg(num_agents=1, steps_per_episode=12))
    env.reset()
    src, dst = adjacent_pair(env, set())
    set_state(env, [src], [dst])
    calls = []

    def assign_adjacent(agent_idx):
        old = env._goals_arr[agent_idx]
        env._goal_owner[int(old[0]), int(old[1])] = env.UNASSIGNED_OWNER
        py, px = map(int, env._positions_arr[agent_idx])
        calls.append((py, px))
        for dy, dx in ((0, 1), (1, 0), (0, -1), (-1, 0)):
            ny, nx = py + dy, px + dx
            if not (0 <= ny < env.grid.shape[0] and 0 <= nx < env.grid.shape[1]):
                continue
            if env.grid[ny, nx] != env.EMPTY_CELL or env._occupancy_owner[ny, nx] != env.UNASSIGNED_OWNER:
                continue
            if env._goal_owner[ny, nx] != env.UNASSIGNED_OWNER:
                continue
            env._goals_arr[agent_idx, :] = [ny, nx]
            env._goal_owner[ny, nx] = agent_idx
            return env._goals_arr[agent_idx]
        raise RuntimeError("no adjacent goal")

    monkeypatch.setattr(env, "_assign_new_goal", assign_adjacent)
    info, done = {}, False
    while not done:
        pos, goal = tuple(map(int, env.positions["agent_0"])), tuple(map(int, env.goals["agent_0"]))
        _o, _r, term, trunc, info = env.step({"agent_0": towards(pos, goal)})
        done = term["__all__"] or trunc["__all__"]
    assert info["__all__"]["goals_reached_total"] == 12.0 > float(env._num_agents)
    assert len(calls) == 12 and calls[0] == dst  # hook saw the agent at its NEW cell (mid-step snapshot)


def test_numpy_backend_lifelong_equals_oracle_with_recorded_ranks():
    """Lifelong + numpy backend: same host RNG stream as the reference => identical to a golden trace."""
    from trace_utils import load_trace

    t = load_trace("m21_n8_lifelong")
    cfg = dict(t["config"])
    env = RM(cfg)
    for ep in range(2):
        obs, _ = env.reset()
        assert np.array_equal(env._starts_arr, t["reset_starts"][ep])
        idx = np.flatnonzero(t["step_ep_index"] == ep)
        for s in idx:
            o, r, term, trunc, info = env.step({f"agent_{i}": int(a) for i, a in enumerate(t["step_actions"][s])})
            assert np.array_equal(env._positions_arr, t["step_positions"][s]), f"positions step {s}"
            assert np.array_equal(env._goals_arr, t["step_goals"][s]), f"goals step {s}"
            assert np.array_equal(np.stack([o[a] for a in env.agents]), t["step_flat_obs"][s]), f"obs step {s}"
            got = np.array([info["__all__"][k] for k in refcompat_info_keys()], np.float64)
            assert np.array_equal(got, t["step_info_all"][s]), f"info step {s}"


def refcompat_info_keys():
    from trace_utils import INFO_KEYS

    return INFO_KEYS


# ------------------------------------------------------------------ lock metrics known answers
def lock_env(**kw):
    cfg = base_cfg(env_name="ReferenceModel-1-3", num_agents=2, steps_per_episode=50, sensor_range=1,
                   deadlock_window_steps=2, livelock_window_steps=4, lock_nearby_manhattan=2,
                   lock_progress_epsilon=1, lock_min_neighbors=1)
    cfg.update(kw)
    env = RM(cfg)
    env.reset()
    return env


def test_deadlock_fires_at_step_two_for_on_goal_blocker():
    env = lock_env()
    set_state(env, [(2, 0), (2, 1)], [(2, 2), (2, 1)])
    env._reset_lock_tracking()
    infos = [env.step({"agent_0": RIGHT, "agent_1": NO_OP})[4]["__all__"] for _ in range(3)]
    assert infos[0]["deadlock_event_step"] == 0.0
    assert infos[1]["deadlock_step"] == 1.0 and infos[1]["deadlock_event_step"] == 1.0
    assert infos[1]["livelock_step"] == 0.0 and infos[1]["deadlock_events_total"] == 1.0
    assert infos[2]["deadlock_event_step"] == 0.0 and infos[2]["deadlock_steps_total"] == 2.0


def test_deadlock_uses_current_state_not_sticky_flags():
    env = lock_env()
    set_state(env, [(2, 0), (2, 2)], [(2, 1), (4, 2)])
    env._reset_lock_tracking()
    env.step({"agent_0": RIGHT, "agent_1": NO_OP})
    assert env.goal_reached_once["agent_0"] and bool(env._completed_once_arr[0])
    env.step({"agent_0": LEFT, "agent_1": LEFT})
    assert tuple(map(int, env.positions["agent_0"])) == (2, 0)
    env.step({"agent_0": RIGHT, "agent_1": NO_OP})
    info = env.step({"agent_0": RIGHT, "agent_1": NO_OP})[4]["__all__"]
    assert info["deadlock_step"] == 1.0 and info["deadlock_event_step"] == 1.0


# ------------------------------------------------------------------ dtype / bounds matrix
@pytest.mark.parametrize("norm", [True, False])
@pytest.mark.parametrize("gdist", [True, False])
@pytest.mark.parametrize("mask", [True, False])
@pytest.mark.parametrize("bp", [True, False])
def test_observations_are_float32_and_inside_the_declared_space(norm, gdist, mask, bp):
    env = RM(base_cfg(normalize_goal_delta=norm, include_goal_distance=gdist, include_action_mask_in_obs=mask,
                      include_blocking_pressure_in_obs=bp, validate_observation_space=True, steps_per_episode=12))
    rng = np.random.default_rng(3)
    obs, _ = env.reset()
    D = 25 + 2 + int(gdist) + int(bp) + 5 * int(mask)
    for _ in range(14):
        for a in env.agents:
            assert obs[a].dtype == np.float32 and obs[a].shape == (D,) and env.observation_space.contains(obs[a])
        obs, _r, term, trunc, _i = env.step({a: int(rng.integers(0, 5)) for a in env.agents})
        if term["__all__"] or trunc["__all__"]:
            obs, _ = env.reset()


def test_callback_facing_attributes():
    """Attributes src/trainers/callbacks.py:114-127,266-308 and main.py:154,265,290,310-315 read."""
    env = RM(base_cfg(deterministic=True, steps_per_episode=20))
    env.reset()
    rng = np.random.default_rng(0)
    for _ in range(20):
        _o, _r, term, trunc, info = env.step({a: int(rng.integers(0, 5)) for a in env.agents})
    assert term["__all__"] and trunc["__all__"] and env.step_count == 20
    for name in ("lifelong_mapf", "step_count", "_completed_once_arr", "goal_reached_once",
                 "_episode_goals_reached_total", "_episode_blocking_count", "_episode_deadlock_events",
                 "_episode_livelock_events", "_episode_deadlock_steps", "_episode_livelock_steps", "seed", "grid",
                 "positions", "starts", "goals", "agents", "possible_agents", "observation_spaces", "action_spaces"):
        assert hasattr(env, name), name
    assert env._episode_deadlock_steps == info["__all__"]["deadlock_steps_total"]
    assert env._episode_livelock_steps == info["__all__"]["livelock_steps_total"]
    assert isinstance(env.render(mode="ansi"), str)
    env.close()
