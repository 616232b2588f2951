"""Single-agent (CTE) view on the GPU: the CUDA kernel replays the LIVE-reference traces bit-exactly (flat float32
obs, float64 scalar reward, done flags, info, positions), agrees with the C oracle on random batches, and the
gym-style B = 1 wrapper reproduces a seeded reference run including the numpy layout draws."""
from pathlib import Path

import numpy as np
import pytest

from test_cte_oracle_golden import GOLDEN, load_cfg

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("path", GOLDEN, ids=[p.stem for p in GOLDEN])
def test_cte_kernel_replays_reference_trace(path):
    from dl_reference_models_b200.single_agent import BatchedCteEnv

    z = np.load(path)
    cfg = load_cfg(z)
    B = 3   # the same trace in three envs of one batch
    env = BatchedCteEnv(cfg, B)
    rep = lambda a: np.broadcast_to(a, (B,) + a.shape).copy()  # noqa: E731
    ep = 0
    obs = env.reset(starts=rep(z["reset_starts"][0]), goals=rep(z["reset_goals"][0]))
    assert np.array_equal(obs.cpu().numpy(), rep(z["reset_obs"][0]))
    for t in range(len(z["actions"])):
        if z["reset_before"][t]:
            ep += 1
            obs = env.reset(starts=rep(z["reset_starts"][ep]), goals=rep(z["reset_goals"][ep]))
            assert np.array_equal(obs.cpu().numpy(), rep(z["reset_obs"][ep])), f"reset {ep}"
        obs, reward, term, trunc, info = env.step(rep(z["actions"][t]))
        assert np.array_equal(obs.cpu().numpy(), rep(z["obs"][t])), f"obs, step {t}"
        assert np.array_equal(reward.cpu().numpy(), np.full(B, z["reward"][t])), f"reward, step {t}"
        assert (term.cpu().numpy() == z["terminated"][t]).all() and (trunc.cpu().numpy() == z["truncated"][t]).all()
        assert np.array_equal(info.cpu().numpy(), rep(z["info"][t])), f"info, step {t}"
        assert np.array_equal(env.positions.cpu().numpy(), rep(z["positions"][t])), f"positions, step {t}"
    assert env.poll_errors() == 0


def test_cte_kernel_equals_oracle_on_random_batches():
    from dl_reference_models_b200 import maps
    from dl_reference_models_b200.single_agent import BatchedCteEnv
    from oracle.cte_oracle import CteOracleEnv

    grid = maps.random_obstacle_grid(14, 19, 0.2, 5, min_free=40)
    cfg = {"grid": grid, "num_agents": 11, "steps_per_episode": 25, "seed": 77, "blocking_penalty": -0.125,
           "move_after_goal_penalty": -0.03}
    B = 200
    env = BatchedCteEnv(cfg, B)
    env.reset()
    st, gl = env.starts.cpu().numpy(), env.goals.cpu().numpy()
    orcs = [CteOracleEnv(cfg, grid) for _ in range(B)]
    ref_obs = np.stack([o.reset(st[e], gl[e]) for e, o in enumerate(orcs)])
    assert np.array_equal(env.flat_obs.cpu().numpy(), ref_obs)
    rng = np.random.default_rng(1)
    for t in range(60):
        acts = rng.integers(0, 5, (B, 11)).astype(np.int8)
        obs, reward, term, trunc, info = env.step(acts)
        outs = [o.step(acts[e]) for e, o in enumerate(orcs)]
        assert np.array_equal(obs.cpu().numpy(), np.stack([o[0] for o in outs])), f"step {t}"
        assert np.array_equal(reward.cpu().numpy(), np.array([o[1] for o in outs])), f"reward step {t}"
        assert np.array_equal(term.cpu().numpy().astype(bool), np.array([o[2] for o in outs]))
        assert np.array_equal(trunc.cpu().numpy().astype(bool), np.array([o[3] for o in outs]))
        assert np.array_equal(info.cpu().numpy(), np.stack([o[4] for o in outs]))
        done = (term.cpu().numpy() | trunc.cpu().numpy()).astype(np.uint8)
        if done.any():
            env.reset(mask=done)
            st, gl = env.starts.cpu().numpy(), env.goals.cpu().numpy()
            for e in np.flatnonzero(done):
                ref = orcs[e].reset(st[e], gl[e])
                assert np.array_equal(env.flat_obs[e].cpu().numpy(), ref)


def test_cte_wrapper_reproduces_seeded_reference_run():
    """gym-style wrapper, seeded like the recorded reference env: the numpy layout draws (CTE:155-185) and every
    payload match the live-reference trace without feeding it the layouts."""
    from dl_reference_models_b200.single_agent import ReferenceModel

    z = np.load([p for p in GOLDEN if p.stem == "cte_m21_n4_rand_random"][0])
    cfg = load_cfg(z)
    env = ReferenceModel(cfg)
    obs, info = env.reset()
    assert obs.dtype == np.float32 and np.array_equal(obs, z["reset_obs"][0])
    ep = 0
    for t in range(len(z["actions"])):
        if z["reset_before"][t]:
            ep += 1
            obs, info = env.reset()
            assert np.array_equal(obs, z["reset_obs"][ep]), f"reset {ep}"
        obs, reward, term, trunc, info = env.step(z["actions"][t])
        assert np.array_equal(obs, z["obs"][t]) and isinstance(reward, float) and reward == float(z["reward"][t])
        assert term is bool(z["terminated"][t]) and trunc is bool(z["truncated"][t])
        assert info["action_mask"].dtype == np.int8 and np.array_equal(info["action_mask"], z["obs"][t][-20:].astype(np.int8))
        assert [info[k] for k in ("blocking_count_step", "goals_reached_step", "goals_reached_total", "blocking_count_total")] == list(z["info"][t])
    with pytest.raises(ValueError):
        env.step([0, 0, 0, 7])


def test_cte_device_layout_draws_and_auto_reset_at_batch_scale():
    """B > 1 defaults to the device RNG backend: layouts come from the reset kernel's Philox draw (valid, different per
    env, redrawn at every reset) and ``step(..., auto_reset=True)`` resets finished envs behind the step without a host
    round trip; a sample of the batch is checked against the C oracle across several episodes."""
    import torch

    from dl_reference_models_b200 import maps
    from dl_reference_models_b200.single_agent import BatchedCteEnv
    from oracle.cte_oracle import CteOracleEnv

    grid = maps.get_grid("ReferenceModel-2-1")
    cfg = {"grid": grid, "num_agents": 4, "steps_per_episode": 12, "seed": 5}
    B, S = 4096, 48
    env = BatchedCteEnv(cfg, B)
    assert env.rng_backend == "device"
    env.reset()

    def check_layouts():
        st, gl = env.starts.cpu().numpy().astype(np.int64), env.goals.cpu().numpy().astype(np.int64)
        assert (grid[st[..., 0], st[..., 1]] == 0).all() and (grid[gl[..., 0], gl[..., 1]] == 0).all()
        cells = np.concatenate([st[..., 0] * grid.shape[1] + st[..., 1], gl[..., 0] * grid.shape[1] + gl[..., 1]], axis=1)
        assert all(len(set(r)) == len(r) for r in cells.tolist()), "2N distinct cells per env"
        return st, gl

    st, gl = check_layouts()
    assert len({tuple(r) for r in st.reshape(B, -1).tolist()}) > B // 2, "layouts differ between envs"
    orcs = [CteOracleEnv(cfg, grid) for _ in range(S)]
    ref = np.stack([o.reset(st[e].astype(np.int16), gl[e].astype(np.int16)) for e, o in enumerate(orcs)])
    assert np.array_equal(env.flat_obs[:S].cpu().numpy(), ref)
    gen = torch.Generator().manual_seed(3)
    episodes = 0
    for t in range(40):
        acts = torch.randint(0, 5, (B, 4), dtype=torch.int8, generator=gen)
        obs, reward, term, trunc, info = env.step(acts, auto_reset=True)
        a = acts.numpy()
        outs = [o.step(a[e]) for e, o in enumerate(orcs)]
        done = (term.cpu().numpy() | trunc.cpu().numpy()).astype(bool)
        assert np.array_equal(reward[:S].cpu().numpy(), np.array([o[1] for o in outs])), f"reward step {t}"
        assert np.array_equal(done[:S], np.array([o[2] or o[3] for o in outs])), f"done step {t}"
        assert np.array_equal(info[:S].cpu().numpy(), np.stack([o[4] for o in outs]))
        if done.any():
            episodes += int(done.sum())
            st, gl = check_layouts()
        got = obs[:S].cpu().numpy()
        for e, o in enumerate(orcs):
            want = o.reset(st[e].astype(np.int16), gl[e].astype(np.int16)) if done[e] else outs[e][0]
            assert np.array_equal(got[e], want), f"obs env {e} step {t}"
    assert episodes >= 3 * B and env.poll_errors() == 0
    # the numpy backend stays available for batches (seed-compatible env 0)
    env2 = BatchedCteEnv(dict(cfg, rng_backend="numpy"), 3)
    env2.reset()
    assert env2.rng_backend == "numpy" and len(env2._rngs) == 3
