"""ctypes front-end of the CPU oracle (TEST INFRASTRUCTURE ONLY).

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module.  It wraps ``oracle/_build/libmapf_oracle.so`` (built from
``mapf_oracle.c`` by ``oracle/Makefile``), a literal C restatement of
``/root/reference/src/environments/reference_model_multi_agent.py`` (cited as ENV:line).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_SO = _HERE / "_build" / "libmapf_oracle.so"

INFO_KEYS = (
    "goals_reached_step",
    "goals_reached_total",
    "blocking_count_step",
    "blocking_count_total",
    "deadlock_step",
    "livelock_step",
    "deadlock_event_step",
    "livelock_event_step",
    "deadlock_events_total",
    "livelock_events_total",
    "deadlock_steps_total",
    "livelock_steps_total",
    "completion_ratio",
    "throughput",
)

ERR_INVALID_ACTION, ERR_NO_GOAL_CELL, ERR_TOO_FEW_CELLS, ERR_BAD_ARG = -1, -2, -3, -4


class _Config(C.Structure):
    _fields_ = [
        ("rows", C.c_int32),
        ("cols", C.c_int32),
        ("num_agents", C.c_int32),
        ("sensor_range", C.c_int32),
        ("steps_per_episode", C.c_int32),
        ("lifelong_mapf", C.c_int32),
        ("enable_lock_metrics", C.c_int32),
        ("deadlock_window_steps", C.c_int32),
        ("livelock_window_steps", C.c_int32),
        ("lock_nearby_manhattan", C.c_int32),
        ("lock_min_neighbors", C.c_int32),
        ("lock_progress_epsilon", C.c_double),
        ("normalize_goal_delta", C.c_int32),
    ]


class _Outputs(C.Structure):
    _fields_ = [
        ("local_obs", C.c_void_p),
        ("action_mask", C.c_void_p),
        ("goal_delta", C.c_void_p),
        ("goal_distance", C.c_void_p),
        ("blocking_prev", C.c_void_p),
        ("reward", C.c_void_p),
        ("terminated", C.c_void_p),
        ("truncated", C.c_void_p),
        ("blocking", C.c_void_p),
        ("goal_reached_step", C.c_void_p),
        ("info_all", C.c_void_p),
        ("moved", C.c_void_p),
        ("failed_move", C.c_void_p),
        ("intended_next", C.c_void_p),
        ("goal_reassigned", C.c_void_p),
    ]


def build(force: bool = False) -> Path:
    """Compile the oracle with gcc (a few hundred ms)."""
    srcs = [_HERE / "mapf_oracle.c", _HERE / "mapf_oracle.h", _HERE / "cte_oracle.c"]
    if (
        force
        or not _SO.exists()
        or _SO.stat().st_mtime < max(f.stat().st_mtime for f in srcs)
    ):
        subprocess.run(["make", "-C", str(_HERE), "-s", "-B"], check=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(_SO))
        L.oracle_create.restype = C.c_void_p
        L.oracle_create.argtypes = [C.POINTER(_Config), C.c_void_p, C.c_uint64]
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_num_free.argtypes = [C.c_void_p]
        L.oracle_get_free_positions.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_set_layout.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_set_state.argtypes = [C.c_void_p] + [C.c_void_p] * 8
        L.oracle_reset_lock_tracking.argtypes = [C.c_void_p]
        L.oracle_reset.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(_Outputs)]
        L.oracle_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(_Outputs)]
        L.oracle_get_state.argtypes = [C.c_void_p] + [C.c_void_p] * 8
        L.oracle_get_owner_grids.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_get_last_candidate_counts.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_get_last_ranks.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_reset_many.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.POINTER(_Outputs)]
        L.oracle_step_many.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.POINTER(_Outputs), C.c_void_p]
        L.oracle_get_state_many.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 8
        L.oracle_set_philox.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_int64, C.c_uint32]
        L.oracle_get_philox_counter.restype = C.c_uint32
        L.oracle_get_philox_counter.argtypes = [C.c_void_p]
        L.oracle_sample_actions_philox.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p]
        L.oracle_sample_actions_philox_many.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p]
        L.oracle_philox_raw.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p]
        L.oracle_bench_run.restype = C.c_int64
        L.oracle_bench_run.argtypes = [
            C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int,
            C.POINTER(C.c_int64), C.POINTER(C.c_uint64),
        ]
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OracleError(Exception):
    def __init__(self, code):
        super().__init__(f"oracle error {code}")
        self.code = code


def make_config(env_config: dict, grid: np.ndarray) -> _Config:
    """env_config uses the reference's keys and defaults (ENV:38-61)."""
    g = env_config.get
    return _Config(
        rows=int(grid.shape[0]),
        cols=int(grid.shape[1]),
        num_agents=int(g("num_agents", 2)),
        sensor_range=int(g("sensor_range", 1)),
        steps_per_episode=int(g("steps_per_episode", 100)),
        lifelong_mapf=int(bool(g("lifelong_mapf", False))),
        enable_lock_metrics=int(bool(g("enable_lock_metrics", True))),
        deadlock_window_steps=max(1, int(g("deadlock_window_steps", 8))),
        livelock_window_steps=max(1, int(g("livelock_window_steps", 16))),
        lock_nearby_manhattan=max(1, int(g("lock_nearby_manhattan", 2))),
        lock_min_neighbors=max(1, int(g("lock_min_neighbors", 1))),
        lock_progress_epsilon=float(g("lock_progress_epsilon", 1)),
        normalize_goal_delta=int(bool(g("normalize_goal_delta", True))),
    )


class StepResult:
    """numpy views of one env's outputs (copied per call)."""

    __slots__ = (
        "local_obs", "action_mask", "goal_delta", "goal_distance", "blocking_prev", "reward",
        "terminated", "truncated", "blocking", "goal_reached_step", "info_all", "moved",
        "failed_move", "intended_next", "goal_reassigned",
    )

    def info_dict(self, lifelong: bool) -> dict:
        n = 14 if lifelong else 12
        return {k: float(self.info_all[i]) for i, k in enumerate(INFO_KEYS[:n])}


class OracleEnv:
    """One reference-equivalent env on the CPU."""

    def __init__(self, env_config: dict, grid: np.ndarray, seed: int = 0):
        self.grid = np.ascontiguousarray(grid, dtype=np.uint8)
        self.cfg = make_config(env_config, self.grid)
        self.N = self.cfg.num_agents
        self.V = 2 * self.cfg.sensor_range + 1
        self.lifelong = bool(self.cfg.lifelong_mapf)
        self._h = lib().oracle_create(C.byref(self.cfg), _ptr(self.grid), C.c_uint64(seed))
        if not self._h:
            raise ValueError("oracle_create failed")
        N, V = self.N, self.V
        self._buf = {
            "local_obs": np.zeros((N, V, V), np.uint8),
            "action_mask": np.zeros((N, 5), np.int8),
            "goal_delta": np.zeros((N, 2), np.float32),
            "goal_distance": np.zeros(N, np.float32),
            "blocking_prev": np.zeros(N, np.float32),
            "reward": np.zeros(N, np.float32),
            "terminated": np.zeros(1, np.uint8),
            "truncated": np.zeros(1, np.uint8),
            "blocking": np.zeros(N, np.float32),
            "goal_reached_step": np.zeros(N, np.float32),
            "info_all": np.zeros(len(INFO_KEYS), np.float64),
            "moved": np.zeros(N, np.uint8),
            "failed_move": np.zeros(N, np.uint8),
            "intended_next": np.zeros((N, 2), np.int16),
            "goal_reassigned": np.zeros(1, np.uint8),
        }
        self._out = _Outputs(**{k: _ptr(v) for k, v in self._buf.items()})

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib is not None:
            _lib.oracle_destroy(h)

    # -- helpers
    def _result(self) -> StepResult:
        r = StepResult()
        for k, v in self._buf.items():
            setattr(r, k, v.copy())
        return r

    @staticmethod
    def _check(rc):
        if rc != 0:
            raise OracleError(rc)

    @property
    def handle(self):
        return self._h

    def free_positions(self) -> np.ndarray:
        f = lib().oracle_num_free(self._h)
        out = np.zeros((f, 2), np.int16)
        lib().oracle_get_free_positions(self._h, _ptr(out))
        return out

    def set_layout(self, starts, goals):
        s = np.ascontiguousarray(starts, np.int16)
        g = np.ascontiguousarray(goals, np.int16)
        self._check(lib().oracle_set_layout(self._h, _ptr(s), _ptr(g)))

    def set_state(self, positions=None, starts=None, goals=None, reached=None, completed_once=None,
                  blocking_prev=None, step_count=None, episode_goals_total=None):
        def a(x, dt):
            return None if x is None else np.ascontiguousarray(x, dt)
        keep = [a(positions, np.int16), a(starts, np.int16), a(goals, np.int16), a(reached, np.uint8),
                a(completed_once, np.uint8), a(blocking_prev, np.float32),
                None if step_count is None else np.array([step_count], np.int32),
                None if episode_goals_total is None else np.array([episode_goals_total], np.float64)]
        self._check(lib().oracle_set_state(self._h, *[_ptr(k) for k in keep]))

    def reset_lock_tracking(self):
        lib().oracle_reset_lock_tracking(self._h)

    def reset(self, mode: int = 2, starts=None, goals=None) -> StepResult:
        s = None if starts is None else np.ascontiguousarray(starts, np.int16)
        g = None if goals is None else np.ascontiguousarray(goals, np.int16)
        self._check(lib().oracle_reset(self._h, mode, _ptr(s), _ptr(g), C.byref(self._out)))
        return self._result()

    def step(self, actions, goal_rank=None, goal_override=None) -> StepResult:
        a = np.ascontiguousarray(actions, np.int8)
        gr = None if goal_rank is None else np.ascontiguousarray(goal_rank, np.int32)
        go = None if goal_override is None else np.ascontiguousarray(goal_override, np.int16)
        self._check(lib().oracle_step(self._h, _ptr(a), _ptr(gr), _ptr(go), C.byref(self._out)))
        return self._result()

    def state(self) -> dict:
        N = self.N
        st = {
            "positions": np.zeros((N, 2), np.int16), "starts": np.zeros((N, 2), np.int16),
            "goals": np.zeros((N, 2), np.int16), "reached": np.zeros(N, np.uint8),
            "completed_once": np.zeros(N, np.uint8), "blocking_prev": np.zeros(N, np.float32),
            "step_count": np.zeros(1, np.int32), "episode_counters": np.zeros(6, np.float64),
        }
        lib().oracle_get_state(self._h, *[_ptr(st[k]) for k in (
            "positions", "starts", "goals", "reached", "completed_once", "blocking_prev",
            "step_count", "episode_counters")])
        return st

    def owner_grids(self):
        occ = np.zeros(self.grid.shape, np.int16)
        goal = np.zeros(self.grid.shape, np.int16)
        lib().oracle_get_owner_grids(self._h, _ptr(occ), _ptr(goal))
        return occ, goal

    def last_candidate_counts(self) -> np.ndarray:
        out = np.zeros(self.N, np.int32)
        lib().oracle_get_last_candidate_counts(self._h, _ptr(out))
        return out


class OracleBatch:
    """B independent oracle envs stepped by one C call; outputs are [B, ...] numpy arrays that are
    overwritten by the next call (same channel names as :class:`StepResult`, plus ``ranks``)."""

    def __init__(self, env_config: dict, grid: np.ndarray, num_envs: int, seed: int = 0):
        grid = np.ascontiguousarray(grid, np.uint8)
        per_env = grid.ndim == 3
        self.envs = [OracleEnv(env_config, grid[b] if per_env else grid, seed=seed + b)
                     for b in range(num_envs)]
        self.B, self.N, self.V = num_envs, self.envs[0].N, self.envs[0].V
        self.lifelong = self.envs[0].lifelong
        self._arr = (C.c_void_p * num_envs)(*[e.handle for e in self.envs])
        B, N, V = self.B, self.N, self.V
        self.buf = {
            "local_obs": np.zeros((B, N, V, V), np.uint8), "action_mask": np.zeros((B, N, 5), np.int8),
            "goal_delta": np.zeros((B, N, 2), np.float32), "goal_distance": np.zeros((B, N), np.float32),
            "blocking_prev": np.zeros((B, N), np.float32), "reward": np.zeros((B, N), np.float32),
            "terminated": np.zeros(B, np.uint8), "truncated": np.zeros(B, np.uint8),
            "blocking": np.zeros((B, N), np.float32), "goal_reached_step": np.zeros((B, N), np.float32),
            "info_all": np.zeros((B, len(INFO_KEYS)), np.float64), "moved": np.zeros((B, N), np.uint8),
            "failed_move": np.zeros((B, N), np.uint8), "intended_next": np.zeros((B, N, 2), np.int16),
            "goal_reassigned": np.zeros(B, np.uint8),
        }
        self.ranks = np.full((B, N), -1, np.int32)
        self._out = _Outputs(**{k: _ptr(v) for k, v in self.buf.items()})

    def reset(self, mode: int = 2, starts=None, goals=None, mask=None) -> dict:
        s = None if starts is None else np.ascontiguousarray(starts, np.int16)
        g = None if goals is None else np.ascontiguousarray(goals, np.int16)
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        OracleEnv._check(lib().oracle_reset_many(self._arr, self.B, mode, _ptr(s), _ptr(g), _ptr(m),
                                                 C.byref(self._out)))
        return self.buf

    def step(self, actions, goal_rank=None, goal_override=None) -> dict:
        a = np.ascontiguousarray(actions, np.int8)
        gr = None if goal_rank is None else np.ascontiguousarray(goal_rank, np.int32)
        go = None if goal_override is None else np.ascontiguousarray(goal_override, np.int16)
        OracleEnv._check(lib().oracle_step_many(self._arr, self.B, _ptr(a), _ptr(gr), _ptr(go),
                                                C.byref(self._out), _ptr(self.ranks)))
        return self.buf

    # -- device-RNG replay (Philox draws of the CUDA kernels; no reference counterpart)
    def set_philox(self, seed: int, env_id_base: int = 0, counters=None):
        """Env b draws its layouts / lifelong goals like the kernels do for global env id ``env_id_base + b``."""
        for b, e in enumerate(self.envs):
            lib().oracle_set_philox(e.handle, 1, C.c_uint64(seed), C.c_int64(env_id_base + b),
                                    C.c_uint32(0 if counters is None else int(counters[b])))
        self._ph_base = env_id_base

    def philox_counters(self) -> np.ndarray:
        return np.array([lib().oracle_get_philox_counter(e.handle) for e in self.envs], np.uint32)

    def sample_actions(self, call_counter: int, masked: bool = True) -> np.ndarray:
        """The benchmark samplers with the kernels' draws, from the current ``buf['action_mask']``."""
        out = np.zeros((self.B, self.N), np.int8)
        mask = np.ascontiguousarray(self.buf["action_mask"])
        lib().oracle_sample_actions_philox_many(self._arr, self.B, C.c_uint64(call_counter), int(masked),
                                                _ptr(mask), _ptr(out))
        return out

    def state(self) -> dict:
        B, N = self.B, self.N
        st = {
            "positions": np.zeros((B, N, 2), np.int16), "starts": np.zeros((B, N, 2), np.int16),
            "goals": np.zeros((B, N, 2), np.int16), "reached": np.zeros((B, N), np.uint8),
            "completed_once": np.zeros((B, N), np.uint8), "blocking_prev": np.zeros((B, N), np.float32),
            "step_count": np.zeros(B, np.int32), "episode_counters": np.zeros((B, 6), np.float64),
        }
        lib().oracle_get_state_many(self._arr, self.B, *[_ptr(st[k]) for k in (
            "positions", "starts", "goals", "reached", "completed_once", "blocking_prev",
            "step_count", "episode_counters")])
        return st


def philox_raw(k0: int, k1: int, ctr) -> np.ndarray:
    """One Philox4x32-10 block (known-answer tests of the replay RNG)."""
    c = np.array(ctr, np.uint32)
    lib().oracle_philox_raw(C.c_uint32(k0), C.c_uint32(k1), _ptr(c))
    return c


def flat_obs_batch(buf: dict, include_goal_distance=False, include_blocking_pressure=True,
                   include_action_mask=False) -> np.ndarray:
    """ENV:306-328 for [B, N, ...] channel arrays -> [B, N, D] float32."""
    B, N = buf["local_obs"].shape[:2]
    parts = [buf["local_obs"].reshape(B, N, -1).astype(np.float32), buf["goal_delta"].astype(np.float32)]
    if include_goal_distance:
        parts.append(buf["goal_distance"].reshape(B, N, 1))
    if include_blocking_pressure:
        parts.append(buf["blocking_prev"].reshape(B, N, 1))
    if include_action_mask:
        parts.append(buf["action_mask"].astype(np.float32))
    return np.concatenate(parts, axis=2).astype(np.float32)


def flat_obs(res: StepResult, include_goal_distance=False, include_blocking_pressure=True,
             include_action_mask=False) -> np.ndarray:
    """ENV:306-328: [N, D] float32 in the reference's component order."""
    N = res.local_obs.shape[0]
    parts = [res.local_obs.reshape(N, -1).astype(np.float32), res.goal_delta.astype(np.float32)]
    if include_goal_distance:
        parts.append(res.goal_distance.reshape(N, 1))
    if include_blocking_pressure:
        parts.append(res.blocking_prev.reshape(N, 1))
    if include_action_mask:
        parts.append(res.action_mask.astype(np.float32))
    return np.concatenate(parts, axis=1).astype(np.float32)


def bench_envs(env_config: dict, grid: np.ndarray, num_envs: int, deterministic_layout=None) -> list:
    envs = [OracleEnv(env_config, grid, seed=int(env_config.get("seed") or 0) + i) for i in range(num_envs)]
    if deterministic_layout is not None:
        for e in envs:
            e.set_layout(*deterministic_layout)
    return envs


def bench_run(env_config: dict, grid: np.ndarray, num_envs: int, steps: int, mode: str = "random",
              deterministic_layout=None, action_seed: int = 999, threads: int | None = None, envs: list | None = None):
    """Time the oracle's benchmark loop (mirrors scripts/benchmark_multi_agent_env.py:59-107).

    Returns (env_steps, episodes, elapsed_s, threads).  ``envs`` (from :func:`bench_envs`) lets
    repeated samples reuse the same env objects."""
    import time

    threads = threads or len(os.sched_getaffinity(0))
    if envs is None:
        envs = bench_envs(env_config, grid, num_envs, deterministic_layout)
    num_envs = len(envs)
    det = 1 if deterministic_layout is not None else 0
    arr = (C.c_void_p * num_envs)(*[e.handle for e in envs])
    eps = C.c_int64(0)
    cs = C.c_uint64(0)
    t0 = time.perf_counter()
    n = lib().oracle_bench_run(arr, num_envs, steps, 0 if mode == "random" else 1, det,
                               C.c_uint64(action_seed), threads, C.byref(eps), C.byref(cs))
    dt = time.perf_counter() - t0
    return int(n), int(eps.value), dt, min(threads, num_envs)
