"""Golden traces of the reference's single-agent (CTE) env from the LIVE Python reference (build container only).

Run:  python tests/golden/make_golden_cte.py
Needs /root/reference (read-only) plus the import stubs in oracle/ref_stubs.  Writes tests/golden/cte_*.npz.

Each trace records, per step of the unmodified env (src/environments/reference_model_single_agent.py): the joint
action, and every output: the flat float32 observation (full grid + 5N mask), the scalar reward (Python float ->
float64), terminated / truncated, info (action_mask, blocking_count_step, goals_reached_step, goals_reached_total,
blocking_count_total), positions; plus the layout (starts / goals) of every reset.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REPO = HERE.parents[1]
sys.dont_write_bytecode = True
sys.path.insert(0, str(REPO / "oracle" / "ref_stubs"))
sys.path.insert(1, "/root/reference")
sys.path.insert(2, str(REPO))

from src.environments import get_grid as ref_get_grid  # noqa: E402
from src.environments.reference_model_single_agent import ReferenceModel  # noqa: E402

INFO_KEYS = ("blocking_count_step", "goals_reached_step", "goals_reached_total", "blocking_count_total")


def layout(env):
    n = env.num_agents
    return (np.array([np.asarray(env.starts[f"agent_{i}"]) for i in range(n)], np.int16),
            np.array([np.asarray(env.goals[f"agent_{i}"]) for i in range(n)], np.int16))


def record(name, cfg, steps, policy, action_seed, grid=None):
    if grid is not None:
        orig = ref_get_grid.get_grid
        ref_get_grid.get_grid = lambda _n: np.array(grid, dtype=np.uint8)
    try:
        env = ReferenceModel(dict(cfg))
    finally:
        if grid is not None:
            ref_get_grid.get_grid = orig
    n = env.num_agents
    rng = np.random.default_rng(action_seed)
    rec = {k: [] for k in ("actions", "obs", "reward", "terminated", "truncated", "info", "positions", "reset_before",
                           "starts", "goals")}
    obs, info = env.reset()
    reset_obs = [obs.copy()]
    reset_layouts = [layout(env)]
    need_reset = False
    for _ in range(steps):
        rec["reset_before"].append(need_reset)
        if need_reset:
            obs, info = env.reset()
            reset_obs.append(obs.copy())
            reset_layouts.append(layout(env))
            need_reset = False
        mask = np.asarray(info["action_mask"]).reshape(n, 5)
        if policy == "random":
            act = rng.integers(0, 5, n)
        elif policy == "masked":
            act = np.array([rng.choice(np.flatnonzero(mask[i])) for i in range(n)])
        else:  # greedy towards the goal, ties and blocked moves resolved by the mask
            act = np.zeros(n, np.int64)
            for i in range(n):
                p, g = np.asarray(env.positions[f"agent_{i}"]), np.asarray(env.goals[f"agent_{i}"])
                pref = []
                if g[0] < p[0]: pref.append(1)
                if g[1] > p[1]: pref.append(2)
                if g[0] > p[0]: pref.append(3)
                if g[1] < p[1]: pref.append(4)
                pref = [a for a in pref if mask[i, a]] or ([int(rng.integers(0, 5))] if rng.random() < 0.3 else [0])
                act[i] = pref[int(rng.integers(0, len(pref)))]
        st, gl = layout(env)
        rec["starts"].append(st); rec["goals"].append(gl)
        obs, reward, term, trunc, info = env.step(act)
        rec["actions"].append(act.astype(np.int8))
        rec["obs"].append(np.asarray(obs, np.float32))
        rec["reward"].append(np.float64(reward))
        rec["terminated"].append(bool(term)); rec["truncated"].append(bool(trunc))
        rec["info"].append(np.array([info[k] for k in INFO_KEYS], np.float64))
        rec["positions"].append(np.array([np.asarray(env.positions[f"agent_{i}"]) for i in range(n)], np.int16))
        need_reset = bool(term or trunc)
    out = {k: np.array(v) for k, v in rec.items()}
    out["reset_obs"] = np.array(reset_obs)
    out["reset_starts"] = np.array([l[0] for l in reset_layouts])
    out["reset_goals"] = np.array([l[1] for l in reset_layouts])
    out["grid"] = np.asarray(env.grid, np.uint8)
    out["cfg_keys"] = np.array(sorted(cfg.keys()))
    out["cfg_vals"] = np.array([str(cfg[k]) for k in sorted(cfg.keys())])
    np.savez_compressed(HERE / f"cte_{name}.npz", **out)
    print(name, "steps", steps, "episodes", len(reset_obs), "reward sum", float(np.sum(out["reward"])))


if __name__ == "__main__":
    record("m21_n4_det_greedy", {"env_name": "ReferenceModel-2-1", "num_agents": 4, "deterministic": True,
                                 "steps_per_episode": 60, "seed": 3}, 300, "greedy", 11)
    record("m21_n4_rand_random", {"env_name": "ReferenceModel-2-1", "num_agents": 4, "deterministic": False,
                                  "steps_per_episode": 40, "seed": 5}, 400, "random", 12)
    record("m14_n4_rand_masked", {"env_name": "ReferenceModel-1-4", "num_agents": 4, "deterministic": False,
                                  "steps_per_episode": 50, "seed": 7, "blocking_penalty": -0.3,
                                  "move_after_goal_penalty": -0.07}, 400, "masked", 13)
    record("m31_n8_rand_greedy", {"env_name": "ReferenceModel-3-1", "num_agents": 8, "deterministic": False,
                                  "steps_per_episode": 80, "seed": 9}, 500, "greedy", 14)
    record("m12_n2_det_random", {"env_name": "ReferenceModel-1-2", "num_agents": 2, "deterministic": True,
                                 "steps_per_episode": 30, "seed": 1}, 200, "random", 15)
