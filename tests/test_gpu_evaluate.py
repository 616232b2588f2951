"""Batched evaluator (main.py:test_trained_model on a GPU batch) against an episode-by-episode replay on the CPU
oracle with the same actions: per-episode totals, timesteps, per-agent returns, lifelong metrics, success rate and
the occupancy heat-map."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def replay_on_oracle(cfg, grid, starts, goals, actions_log):
    from oracle import oracle as orc

    B, N = starts.shape[0], starts.shape[1]
    R, C = grid.shape
    occ = np.zeros((R, C), np.int64)
    rows = []
    for e in range(B):
        env = orc.OracleEnv(cfg, grid, seed=0)
        env.reset(1, starts=starts[e], goals=goals[e])
        total, per_agent, steps, last = 0.0, np.zeros(N), 0, None
        for acts in actions_log:
            o = env.step(acts[e])
            steps += 1
            total += float(np.sum(o.reward.astype(np.float64)))
            per_agent += o.reward.astype(np.float64)
            for (r, c) in env.state()["positions"]:
                occ[r, c] += 1
            last = o
            if bool(np.all(o.terminated)) or bool(np.all(o.truncated)):
                break
        rows.append((total, per_agent, steps, last))
    return rows, occ


def test_evaluator_matches_oracle_replay(lifelong=False):
    import torch

    from dl_reference_models_b200 import maps
    from dl_reference_models_b200.evaluate import evaluate

    grid = maps.get_grid("ReferenceModel-2-1")
    cfg = {"env_name": "ReferenceModel-2-1", "num_agents": 4, "sensor_range": 2, "steps_per_episode": 40,
           "lifelong_mapf": lifelong, "seed": 11, "rng_backend": "philox"}
    log = []
    gen = torch.Generator().manual_seed(3)

    def policy(env, out):
        a = torch.randint(0, 5, (env.B, env.N), dtype=torch.int8, generator=gen)
        # greedy-ish: mostly follow the mask so that episodes can terminate early
        log.append(a.numpy().copy())
        return a.to(env.device)

    B = 48
    res = evaluate(cfg, B, policy=policy)
    starts = np.array([[[r[f"agent_{i}_start_x"], r[f"agent_{i}_start_y"]] for i in range(4)] for r in res.rows], np.int16)
    goals = np.array([[[r[f"agent_{i}_goal_x"], r[f"agent_{i}_goal_y"]] for i in range(4)] for r in res.rows], np.int16)
    ref, occ = replay_on_oracle(cfg, grid, starts, goals, log)
    for e, (total, per_agent, steps, last) in enumerate(ref):
        row = res.rows[e]
        assert row["timesteps"] == steps, e
        assert abs(row["total_reward"] - total) <= 1e-9, e
        assert np.allclose([row[f"agent_{i}_reward"] for i in range(4)], per_agent, atol=1e-9)
    assert np.array_equal(res.occupancy_grid, occ)
    succ = np.mean([1.0 if (bool(np.all(l.terminated)) and not bool(np.all(l.truncated))) else 0.0 for (_, _, _, l) in ref])
    assert res.success_rate == pytest.approx(succ)
    assert res.average_timesteps == pytest.approx(np.mean([s for (_, _, s, _) in ref]))
    df_cols = list(res.rows[0].keys())
    assert df_cols[:5] == ["episode", "cpu_time", "seed", "total_reward", "timesteps"]


def test_evaluator_lifelong_summary_and_heatmap_totals():
    from dl_reference_models_b200 import maps
    from dl_reference_models_b200.evaluate import evaluate

    cfg = {"grid": maps.random_obstacle_grid(32, 32, 0.3, 2026, min_free=32), "num_agents": 16, "sensor_range": 2,
           "steps_per_episode": 64, "lifelong_mapf": True, "seed": 5}
    res = evaluate(cfg, 512, policy="masked")
    assert all(r["timesteps"] == 64 for r in res.rows)
    assert int(res.occupancy_grid.sum()) == 512 * 16 * 64          # every agent, every step, every episode
    assert (res.occupancy_grid[cfg["grid"] == 1] == 0).all()       # nobody ever stands on an obstacle
    for r in res.rows:
        assert r["throughput"] == pytest.approx(r["goals_reached_total"] / 64)
        assert 0.0 <= r["completion_ratio"] <= 1.0
    assert res.success_rate == pytest.approx(np.mean([r["completion_ratio"] for r in res.rows]))
    assert set(res.lifelong) == {"goals_reached_total", "throughput", "completion_ratio"}
