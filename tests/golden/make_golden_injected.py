"""Golden fixtures that need the LIVE Python reference (build container only; see make_golden.py):

* maps.npz        -- all eight obstacle grids of src/environments/get_grid.py:17-727 and the deterministic
                     start / goal tables (:735-872) for every agent count they define.
* injected.npz    -- states no legal step produces, injected the way the reference's own tests do
                     (tests/test_reference_model_multi_agent_invariants.py:28-38: write _positions_arr, rebuild the
                     owner grids): two / three agents on one cell.  Pins the collision penalty (ENV:658-666) and the
                     owner-grid semantics around it (last index wins, ENV:200-205; a leaving agent clears the cell,
                     ENV:523).  "static" scenarios keep the co-located agents in place (NO_OP), "dynamic" ones let
                     everybody act.

Run:  python tests/golden/make_golden_injected.py
"""
from __future__ import annotations

import logging
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

from make_golden import INFO_KEYS, RecordingRng  # noqa: E402

MAPS = ("ReferenceModel-1-1", "ReferenceModel-1-2", "ReferenceModel-1-3", "ReferenceModel-1-4", "ReferenceModel-2-1",
        "ReferenceModel-2-1-b", "ReferenceModel-2-2", "ReferenceModel-3-1")


def dump_maps():
    out = {}
    for name in MAPS:
        key = name.replace("ReferenceModel-", "m").replace("-", "_")
        out[f"{key}_grid"] = np.asarray(ref_get_grid.get_grid(name), np.uint8)
        for n in range(1, 6):
            for what, fn in (("starts", ref_get_grid.get_start_positions), ("goals", ref_get_grid.get_goal_positions)):
                try:
                    d = fn(name, n)
                except Exception as exc:  # the table does not define this agent count
                    out[f"{key}_{what}_{n}_error"] = np.array(type(exc).__name__)
                    continue
                out[f"{key}_{what}_{n}"] = np.array([d[f"agent_{i}"] for i in range(n)], np.int16)
    np.savez_compressed(HERE / "maps.npz", **out)
    print(f"maps.npz: {len(out)} arrays")


def record_injected(lifelong: bool, scenarios: int, steps: int, seed: int):
    n = 6
    cfg = {"env_name": "ReferenceModel-2-1", "num_agents": n, "sensor_range": 2, "steps_per_episode": 5,
           "lifelong_mapf": lifelong, "seed": 77, "deadlock_window_steps": 2, "livelock_window_steps": 3,
           "info_mode": "full", "include_action_mask_in_obs": True, "render_env": False}
    rng = np.random.default_rng(seed)
    rec = {k: [] for k in ("starts", "goals", "injected_positions", "static", "actions", "goal_rank", "positions",
                           "goals_after", "local_obs", "action_mask", "reward", "terminated", "truncated", "moved",
                           "failed_move", "blocking", "info_all")}
    for sc in range(scenarios):
        env = ReferenceModel(cfg)
        env.rng = RecordingRng(env.rng)
        env.reset()
        k = 2 + sc % 2
        members = np.sort(rng.choice(n, size=k, replace=False))
        host = members[sc % k]
        pos = env._positions_arr.copy()
        pos[members] = pos[host]
        env._positions_arr[:] = pos            # the reference tests' _set_state
        env._rebuild_occupancy_owner()
        static = (sc // 2) % 2 == 0
        per = {k2: [] for k2 in ("actions", "goal_rank", "positions", "goals_after", "local_obs", "action_mask", "reward",
                                 "terminated", "truncated", "moved", "failed_move", "blocking", "info_all")}
        starts, goals = env._starts_arr.copy(), env._goals_arr.copy()
        for _ in range(steps):
            acts = rng.integers(0, 5, size=n).astype(np.int8)
            if static:
                acts[members] = 0
            env.rng.draws.clear()
            obs, rew, term, trunc, info = env.step({a: int(acts[i]) for i, a in enumerate(env.agents)})
            rank = np.full(n, -1, np.int32)
            changed = [i for i in range(n) if info[f"agent_{i}"]["goal_reached_step"] > 0] if lifelong else []
            assert len(changed) == len(env.rng.draws)
            for i, (v, _m) in zip(changed, env.rng.draws):
                rank[i] = v
            per["actions"].append(acts)
            per["goal_rank"].append(rank)
            per["positions"].append(env._positions_arr.copy())
            per["goals_after"].append(env._goals_arr.copy())
            per["local_obs"].append(np.stack([info[a]["local_obs"] for a in env.agents]))
            per["action_mask"].append(np.stack([info[a]["action_mask"] for a in env.agents]))
            per["reward"].append(np.array([rew[a] for a in env.agents], np.float64))
            per["terminated"].append(np.uint8(term["__all__"]))
            per["truncated"].append(np.uint8(trunc["__all__"]))
            per["moved"].append(env._scratch_moved_flags.astype(np.uint8))
            per["failed_move"].append(env._scratch_failed_move_flags.astype(np.uint8))
            per["blocking"].append(np.array([info[a]["blocking"] for a in env.agents], np.float32))
            ia = info["__all__"]
            per["info_all"].append(np.array([ia.get(k2, np.nan) for k2 in INFO_KEYS], np.float64))
        rec["starts"].append(starts)
        rec["goals"].append(goals)
        rec["injected_positions"].append(pos)
        rec["static"].append(np.uint8(static))
        for k2, v in per.items():
            rec[k2].append(np.stack(v))
    return {k: np.stack(v) for k, v in rec.items()}, cfg


def main():
    logging.disable(logging.CRITICAL)   # the reference warns about every co-located pair
    dump_maps()
    out = {}
    for lifelong in (False, True):
        r, cfg = record_injected(lifelong, scenarios=48, steps=4, seed=31 + int(lifelong))
        tag = "lifelong" if lifelong else "episodic"
        out.update({f"{tag}_{k}": v for k, v in r.items()})
        pen = int((r["reward"] <= -1.0).sum())
        print(f"injected {tag}: {r['actions'].shape[0]} scenarios x {r['actions'].shape[1]} steps, {pen} penalised agent-steps")
    out["grid"] = np.asarray(ref_get_grid.get_grid("ReferenceModel-2-1"), np.uint8)
    np.savez_compressed(HERE / "injected.npz", **out)


if __name__ == "__main__":
    main()
