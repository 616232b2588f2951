"""GPU parity against the CPU oracle on seeded batches (bit-exact, rewards 1e-6): every env of
the batch, every channel, every step.  The oracle itself is pinned to the reference by
tests/test_oracle_golden.py and tests/test_oracle_reference_digests.py."""
import numpy as np
import pytest

from gpu_utils import ORACLE_STEP_KEYS, STATE_KEYS, assert_batch_equal, gpu_channels

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["lane", "env"])
def kind(request):
    """Both step kernels face the oracle directly (lane-per-agent / env-per-thread)."""
    return request.param


def _run_parity(cfg, grid, B, steps, seed, policy="random", per_env_grid=False, auto_reset=False):
    import torch

    from dl_reference_models_b200.batched_env import BatchedMapfEnv
    from oracle.oracle import OracleBatch, flat_obs_batch

    cfg = dict(cfg)
    cfg["grid"] = grid
    lifelong = bool(cfg.get("lifelong_mapf", False))
    lock = bool(cfg.get("enable_lock_metrics", True))
    ob = OracleBatch(cfg, grid, B, seed=seed)
    env = BatchedMapfEnv(cfg, num_envs=B, device="cuda:0")
    N = env.N
    rng = np.random.default_rng(seed)

    def do_reset(mask=None):
        ob.reset(2, mask=mask)  # oracle draws the layouts (uniform without replacement)
        st = ob.state()
        out = env.reset(mask=mask, starts=st["starts"], goals=st["goals"])
        return st, out

    st, out = do_reset()
    got = gpu_channels(env, out)
    ref = dict(ob.buf)
    ref.update(st)
    ref["flat_obs"] = flat_obs_batch(ob.buf, True, True, True)
    assert_batch_equal(got, ref, ("positions", "goals", "local_obs", "action_mask", "goal_delta", "flat_obs"),
                       "reset", lifelong, lock)
    episodes = 0
    for s in range(steps):
        if policy == "random":
            acts = rng.integers(0, 5, (B, N)).astype(np.int8)
        else:  # masked-random with a push towards the goal: produces arrivals, blocking and locks
            mask = ob.buf["action_mask"].astype(bool)
            acts = np.zeros((B, N), np.int8)
            stt = ob.state()
            d = stt["goals"].astype(int) - stt["positions"].astype(int)
            pref = np.where(np.abs(d[..., 0]) >= np.abs(d[..., 1]),
                            np.where(d[..., 0] < 0, 1, 3), np.where(d[..., 1] > 0, 2, 4))
            pref = np.where((d == 0).all(-1), 0, pref)
            rnd = rng.integers(0, 5, (B, N))
            use_pref = rng.random((B, N)) < 0.7
            acts = np.where(use_pref, pref, rnd).astype(np.int8)
            if policy == "masked":
                ok = np.take_along_axis(mask, acts[..., None].astype(np.int64), axis=2)[..., 0]
                acts = np.where(ok, acts, 0).astype(np.int8)
        ob.step(acts)
        ranks = ob.ranks.copy()
        out = env.step(torch.from_numpy(acts), goal_rank=torch.from_numpy(ranks) if lifelong else None)
        got = gpu_channels(env, out)
        ref = dict(ob.buf)
        ref.update({k: v for k, v in ob.state().items() if k != "blocking_prev"})
        ref["flat_obs"] = flat_obs_batch(ob.buf, True, True, True)
        assert_batch_equal(got, ref, ORACLE_STEP_KEYS + STATE_KEYS + ("flat_obs",), f"step {s}", lifelong, lock)
        done = (ob.buf["terminated"] | ob.buf["truncated"]).astype(np.uint8)
        if done.any():
            episodes += int(done.sum())
            st, out = do_reset(mask=done)
            got = gpu_channels(env, out)
            ref = dict(ob.buf)
            ref.update(st)
            sel = done.astype(bool)
            for k in ("positions", "goals", "local_obs", "action_mask", "goal_delta"):
                assert np.array_equal(got[k][sel], ref[k][sel]), f"masked reset after step {s}: {k}"
    assert env.poll_errors() == 0
    return episodes


def test_c2_4096_envs_4_agents_reference_map():
    """BASELINE config 2: 4096 envs x 4 agents on the 2-1 map, 3 episodes x 100 steps."""
    from dl_reference_models_b200 import maps

    cfg = {"num_agents": 4, "sensor_range": 2, "steps_per_episode": 100, "seed": 123}
    eps = _run_parity(cfg, maps.get_grid("ReferenceModel-2-1"), 4096, 300, seed=999)
    assert eps >= 3 * 4096 - 4096  # early terminations shift the phase of a few envs


def test_c2_goal_seeking_policy_terminations(kind):
    from dl_reference_models_b200 import maps

    cfg = {"num_agents": 4, "sensor_range": 2, "steps_per_episode": 60, "seed": 5, "step_kernel": kind}
    _run_parity(cfg, maps.get_grid("ReferenceModel-2-1"), 1000, 150, seed=11, policy="greedy")


def test_c3_lifelong_32x32_16_agents(kind):
    """BASELINE config 3 shape at a size the oracle finishes in seconds."""
    from dl_reference_models_b200 import maps

    grid = maps.random_obstacle_grid(32, 32, 0.30, 2026, min_free=32)
    cfg = {"num_agents": 16, "sensor_range": 2, "steps_per_episode": 64, "lifelong_mapf": True, "seed": 1,
           "step_kernel": kind}
    _run_parity(cfg, grid, 777, 160, seed=3, policy="greedy")


def test_c4_corridor_32_agents_lock_metrics(kind):
    """BASELINE config 4: all eight lock keys on deadlock-heavy corridors, 64 envs x 256 steps."""
    from dl_reference_models_b200 import maps

    cfg = {"num_agents": 32, "sensor_range": 2, "steps_per_episode": 256, "seed": 2,
           "deadlock_window_steps": 8, "livelock_window_steps": 16, "step_kernel": kind}
    _run_parity(cfg, maps.corridor_grid(32, 32), 64, 256, seed=17, policy="greedy")
    _run_parity(cfg, maps.corridor_grid(32, 32), 33, 128, seed=19, policy="masked")


@pytest.mark.parametrize("n,sr,name,lifelong", [
    (2, 1, "ReferenceModel-1-3", False), (5, 1, "ReferenceModel-1-2", False),
    (3, 3, "ReferenceModel-2-2", True), (8, 2, "ReferenceModel-2-1", True),
    (13, 3, "ReferenceModel-3-1", False), (32, 3, "ReferenceModel-3-1", True),
    (4, 2, "ReferenceModel-1-4", False), (2, 2, "ReferenceModel-1-1", True),
])
def test_shapes_and_agent_counts(n, sr, name, lifelong, kind):
    from dl_reference_models_b200 import maps

    cfg = {"num_agents": n, "sensor_range": sr, "steps_per_episode": 40, "lifelong_mapf": lifelong, "step_kernel": kind,
           "seed": 7, "deadlock_window_steps": 3, "livelock_window_steps": 5,
           "lock_nearby_manhattan": 3, "lock_min_neighbors": 1 if n < 4 else 2,
           "normalize_goal_delta": n % 2 == 0}
    _run_parity(cfg, maps.get_grid(name), 67, 90, seed=n * 10 + sr, policy="greedy")


def test_lock_metrics_disabled_and_windows_32(kind):
    from dl_reference_models_b200 import maps

    g = maps.get_grid("ReferenceModel-1-4")
    _run_parity({"num_agents": 4, "sensor_range": 2, "steps_per_episode": 50, "enable_lock_metrics": False,
                 "step_kernel": kind}, g, 40, 120, seed=1, policy="greedy")
    _run_parity({"num_agents": 4, "sensor_range": 2, "steps_per_episode": 90, "deadlock_window_steps": 32,
                 "livelock_window_steps": 32, "lock_progress_epsilon": 2.5, "step_kernel": kind}, g, 40, 200, seed=2,
                policy="greedy")
    _run_parity({"num_agents": 4, "sensor_range": 2, "steps_per_episode": 90, "deadlock_window_steps": 1,
                 "livelock_window_steps": 1, "lock_progress_epsilon": 0, "step_kernel": kind}, g, 40, 100, seed=3,
                policy="greedy")


def test_per_env_maps():
    from dl_reference_models_b200 import maps

    B = 50
    grids = np.stack([maps.random_obstacle_grid(12, 17, 0.25, 100 + b, min_free=20) for b in range(B)])
    cfg = {"num_agents": 6, "sensor_range": 2, "steps_per_episode": 30, "lifelong_mapf": True, "seed": 9}
    _run_parity(cfg, grids, B, 70, seed=23, policy="greedy")
