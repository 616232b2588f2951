"""Oracle parity of the configuration bench.py actually times: Philox lifelong goal draws, in-launch
auto-reset with Philox layout draws, and the masked sampler fused into the step launch -- every
channel of every env, every step, both step kernels, the compile-time-specialised instantiations
(`FAST` / `MODE`) and the generic ones.

The reference draws with numpy PCG64 and north_star accepts a different stream, so the draws
themselves have no reference counterpart; the oracle restates the kernels' documented draw
discipline (oracle_set_philox, pinned to the Random123 known-answer vectors in
tests/test_oracle_philox.py) and everything around the draws is the reference's transition."""
import os

import numpy as np
import pytest

from gpu_utils import ORACLE_STEP_KEYS, assert_batch_equal, gpu_channels

pytestmark = pytest.mark.gpu


def _replay(cfg, grid, B, steps, kind, fast, env_id_base=0, masked=True):
    import torch

    from dl_reference_models_b200 import _native as nat
    from dl_reference_models_b200.batched_env import BatchedMapfEnv
    from oracle.oracle import OracleBatch

    old = {k: os.environ.get(k) for k in ("MAPF_ENV_FAST", "MAPF_LANE_FAST")}
    os.environ["MAPF_ENV_FAST"] = os.environ["MAPF_LANE_FAST"] = "1" if fast else "0"
    try:
        env = BatchedMapfEnv(dict(cfg, grid=grid, step_kernel=kind), B, "cuda:0", env_id_base=env_id_base)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    lifelong = bool(cfg.get("lifelong_mapf", False))
    ob = OracleBatch(cfg, grid, B)
    ob.set_philox(int(cfg["seed"]), env_id_base)
    out = env.reset()                       # Philox layout draws on the device
    ob.reset(2)                             # ... and the same draws in the oracle
    st = ob.state()
    got = gpu_channels(env, out, flat=False)
    for k in ("positions", "goals", "starts"):
        assert np.array_equal(got[k], st[k]), f"reset: {k}"
    for k in ("local_obs", "action_mask", "goal_delta"):
        assert np.array_equal(got[k], ob.buf[k]), f"reset: {k}"
    counter = 1
    acts_dev = env.sample_actions(masked=masked)
    oa = ob.sample_actions(counter, masked=masked)
    env.fuse_sampler("masked" if masked else "random")     # every step launch now also draws the next actions
    episodes = 0
    arrivals = 0
    for s in range(steps):
        assert np.array_equal(acts_dev.cpu().numpy(), oa), f"step {s}: sampled actions"
        out = env.step(acts_dev, auto_reset=True)          # ONE launch: step + goal draws + reset + sampler
        ob.step(oa)
        ref = {k: v.copy() for k, v in ob.buf.items()}
        arrivals += int(ref["goal_reached_step"].sum())
        done = (ref["terminated"] | ref["truncated"]).astype(np.uint8)
        if done.any():                                      # run_benchmark's `if done: reset()` (benchmark script :89-95)
            episodes += int(done.sum())
            ob.reset(2, mask=done)
            sel = done.astype(bool)
            for k in ("local_obs", "action_mask", "goal_delta", "blocking_prev"):
                ref[k][sel] = ob.buf[k][sel]
            ob.buf["action_mask"][:] = ref["action_mask"]   # the sampler sees the masks of all envs
        st = ob.state()
        ref.update({k: st[k] for k in ("positions", "goals", "starts", "reached", "completed_once")})
        got = gpu_channels(env, out, flat=False)
        assert_batch_equal(got, ref, ORACLE_STEP_KEYS + ("positions", "goals", "starts", "reached", "completed_once"),
                           f"step {s} ({kind}, fast={fast})", lifelong, True)
        words = env.state["env_words"].cpu().numpy()
        assert np.array_equal(words[:, nat.W_STEP_COUNT], st["step_count"]), f"step {s}: step_count"
        assert np.array_equal(words[:, nat.W_RNG_COUNTER].astype(np.uint32), ob.philox_counters()), f"step {s}: RNG counter"
        counter += 1
        oa = ob.sample_actions(counter, masked=masked)
    assert env.poll_errors() == 0
    return episodes, arrivals


@pytest.mark.parametrize("kind", ["env", "lane", "pair"])
@pytest.mark.parametrize("fast", [True, False])
def test_c3_philox_autoreset_fused_sampler(kind, fast):
    """BASELINE config 3 as benchmarked (32x32, 16 agents, lifelong, lock metrics, masked sampler), 4 096 envs x 320
    steps with 64-step episodes: five in-launch resets per env, thousands of Philox goal draws."""
    from dl_reference_models_b200 import maps

    grid = maps.random_obstacle_grid(32, 32, 0.30, 2026, min_free=32)
    cfg = {"num_agents": 16, "sensor_range": 2, "steps_per_episode": 64, "lifelong_mapf": True, "seed": 999,
           "enable_lock_metrics": True, "deadlock_window_steps": 8, "livelock_window_steps": 16}
    episodes, arrivals = _replay(cfg, grid, 4096, 320, kind, fast, env_id_base=3 * 65536)
    assert episodes == 5 * 4096 and arrivals > 4096


@pytest.mark.parametrize("kind", ["env", "lane"])
def test_c3_bench_episode_length_256(kind):
    """The exact bench.py configuration (256-step episodes) across an episode boundary."""
    from dl_reference_models_b200 import maps

    grid = maps.random_obstacle_grid(32, 32, 0.30, 2026, min_free=32)
    cfg = {"num_agents": 16, "sensor_range": 2, "steps_per_episode": 256, "lifelong_mapf": True, "seed": 999,
           "enable_lock_metrics": True, "deadlock_window_steps": 8, "livelock_window_steps": 16}
    episodes, _ = _replay(cfg, grid, 1024, 300, kind, True)
    assert episodes == 1024


@pytest.mark.parametrize("kind", ["env", "lane"])
def test_c4_corridors_episodic_autoreset(kind):
    """BASELINE config 4 as benchmarked: 32 agents on corridors, episodic, masked sampler, auto-reset."""
    from dl_reference_models_b200 import maps

    cfg = {"num_agents": 32, "sensor_range": 2, "steps_per_episode": 48, "lifelong_mapf": False, "seed": 999,
           "enable_lock_metrics": True, "deadlock_window_steps": 8, "livelock_window_steps": 16}
    episodes, _ = _replay(cfg, maps.corridor_grid(32, 32), 512, 150, kind, True)
    assert episodes >= 3 * 512


@pytest.mark.parametrize("kind", ["env", "lane"])
def test_c2_small_map_unmasked_sampler(kind):
    """C2 shape (10x20 reference map, 4 agents) with the unmasked sampler and early terminations."""
    from dl_reference_models_b200 import maps

    cfg = {"num_agents": 4, "sensor_range": 2, "steps_per_episode": 40, "lifelong_mapf": False, "seed": 7}
    _replay(cfg, maps.get_grid("ReferenceModel-2-1"), 2048, 130, kind, True, env_id_base=17, masked=False)


def test_c3_full_baseline_batch_65536_envs():
    """BASELINE configs[2] at its FULL size -- 65 536 envs x 16 agents, the batch bench.py times, with the kernel "auto"
    picks for it (env-per-thread, FAST) -- compared with the oracle on every channel of every env: 20 steps with
    8-step episodes, so every env is reset twice inside a launch and the Philox goal / layout / action draws of a
    million agents are replayed."""
    from dl_reference_models_b200 import maps

    grid = maps.random_obstacle_grid(32, 32, 0.30, 2026, min_free=32)
    cfg = {"num_agents": 16, "sensor_range": 2, "steps_per_episode": 8, "lifelong_mapf": True, "seed": 999,
           "enable_lock_metrics": True, "deadlock_window_steps": 3, "livelock_window_steps": 6}
    episodes, arrivals = _replay(cfg, grid, 65536, 20, "auto", True)
    assert episodes == 2 * 65536 and arrivals > 10000
