"""Helpers shared by the oracle and GPU parity tests: golden-trace loading and replay."""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np

GOLDEN = Path(__file__).resolve().parent / "golden"

INFO_KEYS = (
    "goals_reached_step", "goals_reached_total", "blocking_count_step", "blocking_count_total",
    "deadlock_step", "livelock_step", "deadlock_event_step", "livelock_event_step",
    "deadlock_events_total", "livelock_events_total", "deadlock_steps_total",
    "livelock_steps_total", "completion_ratio", "throughput",
)


def trace_names():
    return sorted(p.stem[len("trace_"):] for p in GOLDEN.glob("trace_*.npz"))


def load_trace(name: str) -> dict:
    with np.load(GOLDEN / f"trace_{name}.npz") as z:
        t = {k: z[k] for k in z.files}
    t["config"] = json.loads(str(t["config_json"]))
    t["name"] = name
    return t


def episode_slices(t: dict):
    """[(episode, first_step, last_step_exclusive)]"""
    ep = t["step_ep_index"]
    out = []
    for e in np.unique(ep):
        idx = np.flatnonzero(ep == e)
        out.append((int(e), int(idx[0]), int(idx[-1]) + 1))
    return out


def assert_step_matches(t: dict, s: int, got: dict, ctx: str = ""):
    """got: channel name -> numpy array for ONE env at trace step s (bit-exact, rewards 1e-6)."""
    cfg = t["config"]
    lifelong = bool(cfg.get("lifelong_mapf", False))
    lock = bool(cfg.get("enable_lock_metrics", True))
    where = f"{t['name']} step {s} {ctx}"
    exact = ["positions", "goals", "local_obs", "action_mask", "goal_delta", "blocking",
             "goal_reached_step", "intended_next", "reached", "completed_once", "flat_obs"]
    if lock:
        exact += ["moved", "failed_move"]
    for k in exact:
        if k not in got:
            continue
        ref = t[f"step_{k}"][s]
        g = np.asarray(got[k])
        assert g.dtype == ref.dtype, f"{where}: {k} dtype {g.dtype} != {ref.dtype}"
        assert g.shape == ref.shape, f"{where}: {k} shape {g.shape} != {ref.shape}"
        assert np.array_equal(g, ref), f"{where}: {k} mismatch\n got={g}\n ref={ref}"
    if "reward" in got:
        np.testing.assert_allclose(np.asarray(got["reward"], np.float64), t["step_reward"][s],
                                   rtol=0, atol=1e-6, err_msg=f"{where}: reward")
    if "terminated" in got:
        assert int(got["terminated"]) == int(t["step_terminated"][s, -1]), f"{where}: terminated"
        assert np.all(t["step_terminated"][s] == t["step_terminated"][s, -1])
    if "truncated" in got:
        assert int(got["truncated"]) == int(t["step_truncated"][s, -1]), f"{where}: truncated"
        assert np.all(t["step_truncated"][s] == t["step_truncated"][s, -1])
    if "info_all" in got:
        n = 14 if lifelong else 12
        ref = t["step_info_all"][s, :n]
        g = np.asarray(got["info_all"], np.float64)[:n]
        assert np.array_equal(g, ref), f"{where}: info_all mismatch\n got={g}\n ref={ref}\n keys={INFO_KEYS[:n]}"
