"""ctypes front-end of the CPU oracle of the single-agent (CTE) env view (TEST INFRASTRUCTURE ONLY).

Wraps ``cte_oracle.c`` (in ``oracle/_build/libmapf_oracle.so``), a literal C restatement of
``/root/reference/src/environments/reference_model_single_agent.py`` (cited CTE:line).  Only tests/ may import it.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import oracle as _o

INFO_KEYS = ("blocking_count_step", "goals_reached_step", "goals_reached_total", "blocking_count_total")


class _Config(C.Structure):
    _fields_ = [("rows", C.c_int32), ("cols", C.c_int32), ("num_agents", C.c_int32), ("steps_per_episode", C.c_int32),
                ("blocking_penalty", C.c_double), ("move_after_goal_penalty", C.c_double)]


class _State(C.Structure):
    _fields_ = [("positions", C.c_void_p), ("goals", C.c_void_p), ("reached_once", C.c_void_p),
                ("step_count", C.c_void_p), ("blocking_total", C.c_void_p)]


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class CteOracleEnv:
    """One env of the CTE view.  ``cfg``: the reference's env_config keys (CTE:84-93)."""

    def __init__(self, cfg: dict, grid: np.ndarray):
        self.lib = _o.lib()
        self.grid = np.ascontiguousarray(grid, np.uint8)
        R, Cc = self.grid.shape
        self.N = int(cfg.get("num_agents", 2))
        self.cfg = _Config(R, Cc, self.N, int(cfg.get("steps_per_episode", 100)),
                           float(cfg.get("blocking_penalty", -0.2)), float(cfg.get("move_after_goal_penalty", -0.05)))
        self.positions = np.zeros((self.N, 2), np.int16)
        self.goals = np.zeros((self.N, 2), np.int16)
        self.reached_once = np.zeros(self.N, np.uint8)
        self.step_count = np.zeros(1, np.int32)
        self.blocking_total = np.zeros(1, np.float64)
        self._st = _State(_ptr(self.positions), _ptr(self.goals), _ptr(self.reached_once), _ptr(self.step_count),
                          _ptr(self.blocking_total))
        self.obs = np.zeros((R, Cc), np.uint8)
        self.mask = np.zeros(5 * self.N, np.int8)

    def reset(self, starts, goals):
        self.positions[:] = np.asarray(starts, np.int16)
        self.goals[:] = np.asarray(goals, np.int16)
        self.lib.cte_oracle_reset(C.byref(self.cfg), C.byref(self._st))
        self.lib.cte_oracle_get_obs(C.byref(self.cfg), _ptr(self.grid), C.byref(self._st), _ptr(self.obs))
        self.lib.cte_oracle_get_action_mask(C.byref(self.cfg), C.byref(self._st), _ptr(self.obs), _ptr(self.mask))
        return self.flat_obs()

    def flat_obs(self) -> np.ndarray:
        return np.concatenate([self.obs.astype(np.float32).reshape(-1), self.mask.astype(np.float32)])

    def step(self, action):
        a = np.ascontiguousarray(action, np.int8)
        reward = C.c_double()
        term, trunc = C.c_uint8(), C.c_uint8()
        info = np.zeros(4, np.float64)
        self.lib.cte_oracle_step.restype = C.c_int
        rc = self.lib.cte_oracle_step(C.byref(self.cfg), _ptr(self.grid), C.byref(self._st), _ptr(a), _ptr(self.obs),
                                      _ptr(self.mask), C.byref(reward), C.byref(term), C.byref(trunc), _ptr(info))
        if rc == -1:
            raise ValueError("Invalid action")
        return self.flat_obs(), float(reward.value), bool(term.value), bool(trunc.value), info
