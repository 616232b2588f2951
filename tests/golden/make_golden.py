"""Generate golden traces from the LIVE Python reference (build container only).

Run:  python tests/golden/make_golden.py
Needs /root/reference (read-only) plus the import stubs in oracle/ref_stubs (the container has
no gymnasium / ray / matplotlib).  Writes tests/golden/trace_*.npz; those fixtures travel to
the GPU box, this script and the reference do not need to.

Every trace records, for each step of the unmodified reference env
(src/environments/reference_model_multi_agent.py): the inputs (actions, the rng.integers()
draw of every lifelong goal reassignment, the layout of every reset) and all outputs
(flat obs, local obs, action mask, goal delta, rewards, terminated/truncated, per-agent info,
info["__all__"] incl. the eight lock metrics, positions, goals, moved / failed_move /
intended_next scratch flags).
"""
from __future__ import annotations

import json
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
from src.environments.reference_model_multi_agent import ReferenceModel  # noqa: E402

from dl_reference_models_b200 import maps as our_maps  # noqa: E402

INFO_KEYS = (
    "goals_reached_step", "goals_reached_total", "blocking_count_step", "blocking_count_total",
    "deadlock_step", "livelock_step", "deadlock_event_step", "livelock_event_step",
    "deadlock_events_total", "livelock_events_total", "deadlock_steps_total",
    "livelock_steps_total", "completion_ratio", "throughput",
)


class RecordingRng:
    """Forwards to the env's numpy Generator and remembers every integers() draw."""

    def __init__(self, rng):
        self._rng = rng
        self.draws = []

    def integers(self, *a, **k):
        v = self._rng.integers(*a, **k)
        self.draws.append((int(v), int(a[0])))
        return v

    def choice(self, *a, **k):
        return self._rng.choice(*a, **k)


def make_env(cfg: dict, grid):
    """Instantiate the reference env; synthetic maps go in through get_grid."""
    if grid is not None:
        orig = ref_get_grid.get_grid
        ref_get_grid.get_grid = lambda name: np.array(grid, dtype=np.uint8)
        try:
            env = ReferenceModel(cfg)
        finally:
            ref_get_grid.get_grid = orig
    else:
        env = ReferenceModel(cfg)
    env.rng = RecordingRng(env.rng)
    return env


def pick_actions(policy: str, env, flat_obs: dict, rng) -> np.ndarray:
    n = env._num_agents
    if policy == "random":
        return rng.integers(0, 5, size=n).astype(np.int8)
    sl = env._obs_slices["action_mask"]
    acts = np.zeros(n, np.int8)
    for i, aid in enumerate(env.agents):
        mask = flat_obs[aid][sl] > 0.5
        valid = np.flatnonzero(mask)
        if policy == "masked":
            acts[i] = rng.choice(valid)
            continue
        # "greedy": with p=0.75 step towards the goal if the mask allows, else masked-random
        # "stubborn": with p=0.9 push towards the goal even against the mask; sit still on goal
        dr, dc = (env._goals_arr[i].astype(int) - env._positions_arr[i].astype(int))
        pref = []
        if dr < 0: pref.append(1)
        if dc > 0: pref.append(2)
        if dr > 0: pref.append(3)
        if dc < 0: pref.append(4)
        if policy == "stubborn":
            if not pref:
                acts[i] = 0
            elif rng.random() < 0.9:
                acts[i] = pref[0]
            else:
                acts[i] = rng.choice(valid)
            continue
        pref = [a for a in pref if mask[a]]
        if pref and rng.random() < 0.75:
            acts[i] = rng.choice(pref)
        else:
            acts[i] = rng.choice(valid)
    return acts


def record(name: str, cfg: dict, *, grid=None, episodes: int, max_steps: int, policy: str,
           action_seed: int):
    cfg = dict(cfg)
    cfg.update(info_mode="full", include_action_mask_in_obs=True, include_goal_distance=True,
               include_blocking_pressure_in_obs=True, render_env=False)
    env = make_env(cfg, grid)
    n = env._num_agents
    rng = np.random.default_rng(action_seed)
    rec = {k: [] for k in (
        "ep_index", "actions", "goal_rank", "positions", "goals", "flat_obs", "local_obs",
        "action_mask", "goal_delta", "reward", "terminated", "truncated", "blocking",
        "goal_reached_step", "info_all", "moved", "failed_move", "intended_next", "reached",
        "completed_once")}
    resets = {k: [] for k in ("starts", "goals", "flat_obs", "local_obs", "action_mask")}
    lifelong = bool(cfg.get("lifelong_mapf", False))

    for ep in range(episodes):
        obs, infos = env.reset()
        resets["starts"].append(env._starts_arr.copy())
        resets["goals"].append(env._goals_arr.copy())
        resets["flat_obs"].append(np.stack([obs[a] for a in env.agents]))
        resets["local_obs"].append(np.stack([infos[a]["local_obs"] for a in env.agents]))
        resets["action_mask"].append(np.stack([infos[a]["action_mask"] for a in env.agents]))
        for _ in range(max_steps):
            acts = pick_actions(policy, env, obs, rng)
            env.rng.draws.clear()
            goals_before = env._goals_arr.copy()
            obs, rew, term, trunc, info = env.step({a: int(acts[i]) for i, a in enumerate(env.agents)})
            # map each integers() draw to the agent whose goal changed, in agent order
            rank = np.full(n, -1, np.int32)
            changed = [i for i in range(n) if info[f"agent_{i}"]["goal_reached_step"] > 0] if lifelong else []
            assert len(changed) == len(env.rng.draws), (changed, env.rng.draws)
            for i, (v, _n) in zip(changed, env.rng.draws):
                rank[i] = v
            del goals_before
            rec["ep_index"].append(ep)
            rec["actions"].append(acts.copy())
            rec["goal_rank"].append(rank)
            rec["positions"].append(env._positions_arr.copy())
            rec["goals"].append(env._goals_arr.copy())
            rec["flat_obs"].append(np.stack([obs[a] for a in env.agents]))
            rec["local_obs"].append(np.stack([info[a]["local_obs"] for a in env.agents]))
            rec["action_mask"].append(np.stack([info[a]["action_mask"] for a in env.agents]))
            rec["goal_delta"].append(np.stack([info[a]["goal_delta"] for a in env.agents]))
            rec["reward"].append(np.array([rew[a] for a in env.agents], np.float64))
            rec["terminated"].append(np.array([term[a] for a in env.agents] + [term["__all__"]], np.uint8))
            rec["truncated"].append(np.array([trunc[a] for a in env.agents] + [trunc["__all__"]], np.uint8))
            rec["blocking"].append(np.array([info[a]["blocking"] for a in env.agents], np.float32))
            rec["goal_reached_step"].append(
                np.array([info[a]["goal_reached_step"] for a in env.agents], np.float32))
            ia = info["__all__"]
            rec["info_all"].append(np.array([ia.get(k, np.nan) for k in INFO_KEYS], np.float64))
            rec["moved"].append(env._scratch_moved_flags.astype(np.uint8))
            rec["failed_move"].append(env._scratch_failed_move_flags.astype(np.uint8))
            rec["intended_next"].append(env._scratch_intended_next.copy())
            rec["reached"].append(env._reached_arr.astype(np.uint8))
            rec["completed_once"].append(env._completed_once_arr.astype(np.uint8))
            if term["__all__"] or trunc["__all__"]:
                break

    out = {f"step_{k}": np.stack(v) for k, v in rec.items()}
    out.update({f"reset_{k}": np.stack(v) for k, v in resets.items()})
    out["grid"] = np.asarray(env.grid, np.uint8)
    out["config_json"] = np.array(json.dumps(cfg))
    out["policy"] = np.array(policy)
    path = HERE / f"trace_{name}.npz"
    np.savez_compressed(path, **out)
    ia = out["step_info_all"]
    print(f"{name:28s} steps={len(rec['actions']):5d} episodes={episodes} "
          f"goals={ia[:, 0].sum():.0f} deadlock_steps={ia[:, 4].sum():.0f} "
          f"livelock_steps={ia[:, 5].sum():.0f} blocking={ia[:, 2].sum():.0f} "
          f"early_term={int(sum(1 for t, u in zip(out['step_terminated'][:, -1], out['step_truncated'][:, -1]) if t and not u))} "
          f"size={path.stat().st_size // 1024} KiB")


def main():
    base = {"seed": 123, "training_execution_mode": "CTDE"}
    # BASELINE config 2 (= the reference's golden-digest config, lock metrics included here)
    record("c2_det_random", {**base, "env_name": "ReferenceModel-2-1", "deterministic": True,
                             "num_agents": 4, "sensor_range": 2, "steps_per_episode": 100},
           episodes=3, max_steps=140, policy="random", action_seed=999)
    record("c2_rand_random", {**base, "env_name": "ReferenceModel-2-1", "deterministic": False,
                              "num_agents": 4, "sensor_range": 2, "steps_per_episode": 100},
           episodes=3, max_steps=140, policy="random", action_seed=999)
    record("c2_rand_greedy", {**base, "env_name": "ReferenceModel-2-1", "deterministic": False,
                              "num_agents": 4, "sensor_range": 2, "steps_per_episode": 100},
           episodes=6, max_steps=140, policy="greedy", action_seed=7)
    # default-ish config: sr=1, 2 agents, tiny map, default windows
    record("m13_n2_sr1", {**base, "env_name": "ReferenceModel-1-3", "deterministic": False,
                          "num_agents": 2, "sensor_range": 1, "steps_per_episode": 60},
           episodes=4, max_steps=60, policy="masked", action_seed=5)
    record("m11_n2_det", {**base, "env_name": "ReferenceModel-1-1", "deterministic": True,
                          "num_agents": 2, "sensor_range": 1, "steps_per_episode": 40},
           episodes=3, max_steps=40, policy="greedy", action_seed=11)
    # lock-heavy: narrow cross map, short windows
    record("m14_n4_lock", {**base, "env_name": "ReferenceModel-1-4", "deterministic": False,
                           "num_agents": 4, "sensor_range": 2, "steps_per_episode": 120,
                           "deadlock_window_steps": 2, "livelock_window_steps": 4},
           episodes=4, max_steps=120, policy="random", action_seed=3)
    record("m12_n5_lock", {**base, "env_name": "ReferenceModel-1-2", "deterministic": False,
                           "num_agents": 5, "sensor_range": 1, "steps_per_episode": 100,
                           "deadlock_window_steps": 3, "livelock_window_steps": 6,
                           "lock_nearby_manhattan": 3, "lock_min_neighbors": 2},
           episodes=3, max_steps=100, policy="greedy", action_seed=21)
    # lifelong
    record("m21_n8_lifelong", {**base, "env_name": "ReferenceModel-2-1", "deterministic": False,
                               "num_agents": 8, "sensor_range": 2, "steps_per_episode": 150,
                               "lifelong_mapf": True},
           episodes=2, max_steps=150, policy="greedy", action_seed=13)
    record("m21_n4_det_lifelong", {**base, "env_name": "ReferenceModel-2-1", "deterministic": True,
                                   "num_agents": 4, "sensor_range": 2, "steps_per_episode": 80,
                                   "lifelong_mapf": True, "normalize_goal_delta": False},
           episodes=3, max_steps=80, policy="greedy", action_seed=17)
    record("m22_n16_sr3_lifelong", {**base, "env_name": "ReferenceModel-2-2", "deterministic": False,
                                    "num_agents": 16, "sensor_range": 3, "steps_per_episode": 120,
                                    "lifelong_mapf": True},
           episodes=2, max_steps=120, policy="greedy", action_seed=19)
    record("m31_n32_sr3_lifelong", {**base, "env_name": "ReferenceModel-3-1", "deterministic": False,
                                    "num_agents": 32, "sensor_range": 3, "steps_per_episode": 100,
                                    "lifelong_mapf": True},
           episodes=1, max_steps=100, policy="greedy", action_seed=23)
    # BASELINE config 3 shape: 32x32 random obstacles, 16 agents, lifelong
    g32 = our_maps.random_obstacle_grid(32, 32, 0.30, 2026, min_free=32)
    record("c3_rand32_n16_lifelong", {**base, "env_name": "synthetic-32x32", "deterministic": False,
                                      "num_agents": 16, "sensor_range": 2, "steps_per_episode": 256,
                                      "lifelong_mapf": True},
           grid=g32, episodes=1, max_steps=256, policy="greedy", action_seed=29)
    record("c3_rand32_n16_random", {**base, "env_name": "synthetic-32x32", "deterministic": False,
                                    "num_agents": 16, "sensor_range": 2, "steps_per_episode": 128,
                                    "lifelong_mapf": True},
           grid=g32, episodes=1, max_steps=128, policy="random", action_seed=31)
    # BASELINE config 4 shape: corridors, 32 agents, non-lifelong, lock parity
    gc = our_maps.corridor_grid(32, 32)
    record("c4_corridor_n32_lock", {**base, "env_name": "synthetic-corridor", "deterministic": False,
                                    "num_agents": 32, "sensor_range": 2, "steps_per_episode": 256,
                                    "deadlock_window_steps": 8, "livelock_window_steps": 16},
           grid=gc, episodes=1, max_steps=256, policy="greedy", action_seed=37)
    record("c4_corridor_n32_random", {**base, "env_name": "synthetic-corridor", "deterministic": False,
                                      "num_agents": 32, "sensor_range": 2, "steps_per_episode": 128,
                                      "deadlock_window_steps": 8, "livelock_window_steps": 16},
           grid=gc, episodes=1, max_steps=128, policy="masked", action_seed=41)
    record("c4_corridor_n32_stubborn", {**base, "env_name": "synthetic-corridor", "deterministic": False,
                                        "num_agents": 32, "sensor_range": 2, "steps_per_episode": 160,
                                        "deadlock_window_steps": 8, "livelock_window_steps": 16},
           grid=gc, episodes=2, max_steps=160, policy="stubborn", action_seed=47)
    record("m21_n4_stubborn", {**base, "env_name": "ReferenceModel-2-1", "deterministic": False,
                               "num_agents": 4, "sensor_range": 2, "steps_per_episode": 60},
           episodes=8, max_steps=60, policy="stubborn", action_seed=53)
    # lock metrics off
    record("m21_n4_nolock", {**base, "env_name": "ReferenceModel-2-1", "deterministic": False,
                             "num_agents": 4, "sensor_range": 2, "steps_per_episode": 50,
                             "enable_lock_metrics": False},
           episodes=2, max_steps=50, policy="greedy", action_seed=43)


if __name__ == "__main__":
    main()
