"""Test-side restatement of the reference's golden-trace hashing protocol
(tests/test_reference_model_multi_agent_parity.py:38-82,101-139 in the reference) and of the
dict payloads ReferenceModel.reset()/step() return (ENV:350-358, 627-656), so that any
implementation exposing per-channel arrays can be checked against the reference's own
SHA-256 digests."""
from __future__ import annotations

import hashlib

import numpy as np

EXPECTED_STOCHASTIC_DIGEST = "d58a9e9e0e383f29c5d7f96a1338dfd73c9035dc5335f66b2ce11a0d8e0452de"
EXPECTED_DETERMINISTIC_DIGEST = "2612dc3eeab5b4fd69d8cbe7fb4f01e35cf52a6f05b07bc955a306ae765c2595"
EXPECTED_STOCHASTIC_SUMMARY = [(0, 100, -3.5), (1, 100, -4.0), (2, 100, -2.5)]
EXPECTED_DETERMINISTIC_SUMMARY = [(0, 100, -3.5), (1, 100, -3.5), (2, 100, -4.0)]

LOCK_KEYS = frozenset({
    "deadlock_step", "livelock_step", "deadlock_event_step", "livelock_event_step",
    "deadlock_events_total", "livelock_events_total", "deadlock_steps_total", "livelock_steps_total",
})


def golden_env_config(deterministic: bool) -> dict:
    return {
        "env_name": "ReferenceModel-2-1", "seed": 123, "deterministic": deterministic,
        "num_agents": 4, "steps_per_episode": 100, "sensor_range": 2, "info_mode": "full",
        "training_execution_mode": "CTDE", "render_env": False,
        "include_action_mask_in_obs": True, "include_blocking_pressure_in_obs": False,
    }


def feed_array(h, arr):
    a = np.asarray(arr)
    h.update(str(a.dtype).encode())
    h.update(str(a.shape).encode())
    h.update(a.tobytes())


def feed(h, v):
    if isinstance(v, dict):
        for k in sorted(v, key=str):
            h.update(str(k).encode())
            feed(h, v[k])
    elif isinstance(v, (list, tuple)):
        for x in v:
            feed(h, x)
    elif isinstance(v, np.ndarray):
        feed_array(h, v)
    elif isinstance(v, (np.floating, float)):
        h.update(np.float32(v).tobytes())
    elif isinstance(v, (np.integer, int, np.bool_, bool)):
        h.update(str(int(v)).encode())
    elif v is None:
        h.update(b"None")
    else:
        h.update(str(v).encode())


def drop_lock(v):
    if isinstance(v, dict):
        return {k: drop_lock(x) for k, x in v.items() if k not in LOCK_KEYS}
    if isinstance(v, list):
        return [drop_lock(x) for x in v]
    if isinstance(v, tuple):
        return tuple(drop_lock(x) for x in v)
    return v


def trace_digest(env, episodes: int = 3, max_steps: int = 140, action_seed: int = 999):
    """Drive any object with the reference's reset()/step() dict API; returns (hex, summary)."""
    rng = np.random.default_rng(action_seed)
    h = hashlib.sha256()
    summary = []
    for ep in range(episodes):
        obs, infos = env.reset()
        h.update(f"episode_{ep}_reset".encode())
        for aid in sorted(obs):
            h.update(aid.encode())
            feed_array(h, obs[aid])
        feed(h, drop_lock(infos))
        total = 0.0
        for s in range(max_steps):
            actions = {f"agent_{i}": int(rng.integers(0, 5)) for i in range(4)}
            obs, rew, term, trunc, infos = env.step(actions)
            total += float(sum(rew.values()))
            h.update(f"episode_{ep}_step_{s}".encode())
            feed(h, actions)
            for aid in sorted(obs):
                h.update(aid.encode())
                feed_array(h, obs[aid])
            feed(h, rew)
            feed(h, term)
            feed(h, trunc)
            feed(h, drop_lock(infos))
            if term.get("__all__", False) or trunc.get("__all__", False):
                summary.append((ep, s + 1, round(total, 6)))
                break
    return h.hexdigest(), summary


class NumpyLayoutStream:
    """The reference's layout RNG stream: default_rng(seed).choice(F, 2N, replace=False), one
    draw in the constructor (ENV:134) and one per reset (ENV:457)."""

    def __init__(self, seed, free_positions: np.ndarray, num_agents: int):
        self.rng = np.random.default_rng(seed)
        self.free = free_positions
        self.n = num_agents

    def draw(self):
        idx = self.rng.choice(self.free.shape[0], size=2 * self.n, replace=False)
        return self.free[idx[: self.n]].copy(), self.free[idx[self.n:]].copy()


class OracleDictEnv:
    """reset()/step() dict API (ENV:440-695) on top of the channel-level CPU oracle."""

    def __init__(self, cfg: dict, grid=None):
        from dl_reference_models_b200 import maps
        from oracle.oracle import OracleEnv

        self.cfg = cfg
        self.grid = maps.get_grid(cfg["env_name"]) if grid is None else grid
        self.o = OracleEnv(cfg, self.grid)
        self.n = int(cfg.get("num_agents", 2))
        self.agents = [f"agent_{i}" for i in range(self.n)]
        self.full = str(cfg.get("info_mode", "lite")).lower() == "full"
        self.lifelong = bool(cfg.get("lifelong_mapf", False))
        self.det = bool(cfg.get("deterministic", False))
        self.flags = (bool(cfg.get("include_goal_distance", False)),
                      bool(cfg.get("include_blocking_pressure_in_obs", True)),
                      bool(cfg.get("include_action_mask_in_obs", False)))
        if self.det:
            s = maps.get_start_positions(cfg["env_name"], self.n)
            g = maps.get_goal_positions(cfg["env_name"], self.n)
            self.o.set_layout([s[a] for a in self.agents], [g[a] for a in self.agents])
        else:
            self.layouts = NumpyLayoutStream(cfg.get("seed"), self.o.free_positions(), self.n)
            self.o.set_layout(*self.layouts.draw())

    def _full_info(self, r, st, i):
        return {"position": st["positions"][i].copy(), "goal": st["goals"][i].copy(),
                "goal_delta": r.goal_delta[i].copy(), "action_mask": r.action_mask[i].copy(),
                "local_obs": r.local_obs[i].copy()}

    def _obs(self, r):
        from oracle.oracle import flat_obs

        f = flat_obs(r, *self.flags)
        return {a: f[i] for i, a in enumerate(self.agents)}

    def reset(self):
        r = self.o.reset(mode=0) if self.det else self.o.reset(1, *self.layouts.draw())
        st = self.o.state()
        infos = {a: (self._full_info(r, st, i) if self.full else {}) for i, a in enumerate(self.agents)}
        return self._obs(r), infos

    def step(self, actions: dict):
        r = self.o.step([actions[a] for a in self.agents])
        st = self.o.state()
        ia = r.info_dict(self.lifelong)
        info = {}
        for i, a in enumerate(self.agents):
            d = self._full_info(r, st, i) if self.full else {}
            d["blocking"] = float(r.blocking[i])
            d["goal_reached_step"] = float(r.goal_reached_step[i])
            d["goals_reached_total"] = ia["goals_reached_total"]
            d["blocking_count_total"] = ia["blocking_count_total"]
            info[a] = d
        info["__all__"] = ia
        rew = {a: float(r.reward[i]) for i, a in enumerate(self.agents)}
        t, u = bool(r.terminated[0]), bool(r.truncated[0])
        term = {a: t for a in self.agents}
        trunc = {a: u for a in self.agents}
        term["__all__"] = t
        trunc["__all__"] = u
        return self._obs(r), rew, term, trunc, info
