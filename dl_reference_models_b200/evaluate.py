"""Batched evaluator: the reference's ``test_trained_model`` (main.py:93-387) with every episode running as one env
of a GPU batch.

The reference plays ``num_episodes`` episodes one after the other, sums the rewards, counts timesteps, tracks
per-agent returns, the start / goal layout, the lifelong metrics of the last ``info["__all__"]`` and an occupancy
heat-map (``occupancy_grid[y, x] += 1`` for every agent after every step), then writes one CSV row per episode.  Here
episode e is env e of a :class:`BatchedMapfEnv`; an env stops contributing at its first ``terminated | truncated``
(no auto-reset), the heat-map is accumulated by ``mapf_occupancy_accumulate`` over the still-active envs, and the
per-episode sums stay on the device until the end.  Policies: ``"random"`` (the reference's ``ALGO_NAME == "RANDOM"``
branch: uniform actions, main.py:212-214), ``"masked"`` (uniform over the action mask) or any callable
``policy(env, out) -> int8 [B,N]`` (e.g. :class:`rollout.ActionMaskPolicy` inference).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import torch

from . import _native as nat
from .batched_env import BatchedMapfEnv


@dataclass
class EvalResult:
    rows: list = field(default_factory=list)          # one dict per episode, the CSV schema of main.py:296-325
    occupancy_grid: np.ndarray | None = None          # int64 [R,C], main.py:153-155
    success_rate: float = 0.0                         # main.py:330
    average_reward: float = 0.0                       # main.py:327
    average_timesteps: float = 0.0                    # main.py:328
    lifelong: dict = field(default_factory=dict)      # main.py:332-338

    def to_dataframe(self):
        import pandas as pd

        return pd.DataFrame(self.rows)

    def save_csv(self, path):
        self.to_dataframe().to_csv(path, index=False)   # main.py:361
        return path


def evaluate(env_config: dict, num_episodes: int, policy="random", device="cuda:0", max_steps: int | None = None) -> EvalResult:
    env = BatchedMapfEnv(env_config, num_episodes, device)
    B, N = env.B, env.N
    dev = env.device
    out = env.reset()
    starts = env.state["starts"].cpu().numpy().copy()
    goals0 = env.state["goals"].cpu().numpy().copy()
    active = torch.ones(B, dtype=torch.bool, device=dev)
    ep_reward = torch.zeros(B, dtype=torch.float64, device=dev)
    agent_reward = torch.zeros((B, N), dtype=torch.float64, device=dev)
    steps = torch.zeros(B, dtype=torch.int64, device=dev)
    success = torch.zeros(B, dtype=torch.float64, device=dev)
    last_info = torch.zeros((B, nat.INFO_WORDS), dtype=torch.int32, device=dev)
    occupancy = torch.zeros(env.grid.shape[-2:], dtype=torch.int64, device=dev)
    limit = max_steps if max_steps is not None else int(env.cfg.steps_per_episode) + 1
    for _ in range(limit):
        if not bool(active.any()):
            break
        if callable(policy):
            actions = policy(env, out)
        else:
            actions = env.sample_actions(masked=(policy == "masked"))
        out = env.step(actions, auto_reset=False)
        act_f = active.to(torch.float64)
        r = out.reward.to(torch.float64)
        ep_reward += r.sum(dim=1) * act_f                      # main.py:251
        agent_reward += r * act_f[:, None]                     # main.py:262
        steps += active.to(torch.int64)
        env.accumulate_occupancy(occupancy, active.to(torch.uint8))   # main.py:264-267
        done = (out.terminated | out.truncated).bool() & active
        last_info[active] = out.info[active]
        success[done] = ((out.terminated != 0) & (out.truncated == 0)).to(torch.float64)[done]   # main.py:294
        active &= ~done
    res = EvalResult()
    info = last_info.cpu().numpy()
    ep_r, ag_r, st = ep_reward.cpu().numpy(), agent_reward.cpu().numpy(), steps.cpu().numpy()
    lifelong = bool(env.lifelong)
    succ = success.cpu().numpy()
    g_tot = info[:, nat.I_GOALS_REACHED_TOTAL].astype(np.float64)
    thr = g_tot / np.maximum(info[:, nat.I_STEP_COUNT], 1)
    comp = info[:, nat.I_COMPLETED_COUNT].astype(np.float64) / float(N)
    for e in range(B):
        row = {"episode": e + 1, "cpu_time": 0.0, "seed": env.env_config.get("seed"), "total_reward": float(ep_r[e]),
               "timesteps": int(st[e])}
        if lifelong:   # main.py:305-315
            row.update(goals_reached_total=float(g_tot[e]), throughput=float(thr[e]), completion_ratio=float(comp[e]))
        for i in range(N):   # main.py:319-325 (x = row index, y = column index in the reference's naming)
            row[f"agent_{i}_reward"] = float(ag_r[e, i])
            row[f"agent_{i}_start_x"], row[f"agent_{i}_start_y"] = int(starts[e, i, 0]), int(starts[e, i, 1])
            row[f"agent_{i}_goal_x"], row[f"agent_{i}_goal_y"] = int(goals0[e, i, 0]), int(goals0[e, i, 1])
        res.rows.append(row)
    res.occupancy_grid = occupancy.cpu().numpy()
    res.success_rate = float(np.mean(comp if lifelong else succ)) if B else 0.0   # main.py:314,316,330
    res.average_reward = float(ep_r.sum() / max(B, 1))
    res.average_timesteps = float(st.sum() / max(B, 1))
    if lifelong:
        res.lifelong = {"goals_reached_total": float(g_tot.mean()), "throughput": float(thr.mean()),
                        "completion_ratio": float(comp.mean())}
    env.close()
    return res
