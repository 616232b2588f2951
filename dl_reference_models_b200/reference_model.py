"""Drop-in ``ReferenceModel``: the reference's RLlib ``MultiAgentEnv`` API on top of the CUDA kernels.

Replaces ``src/environments/reference_model_multi_agent.py`` (``ENV:line`` below) behind the same
plugin hook (``register_env(name, lambda cfg: ReferenceModel(cfg))``, ``main.py:80-82,396``):

* same ``env_config`` keys and defaults (ENV:38-61), same exceptions (ENV:53-55, 270-275, 296-298,
  504-506), same ``reset()`` / ``step()`` signatures and dict payloads (flat float32 obs, Python
  float rewards, Python bool terminated/truncated incl. ``"__all__"``, the info dicts of
  ENV:627-656 and the ``full`` info mode of ENV:350-358);
* the public attributes RLlib, ``main.py`` and the callbacks read, and the private numpy arrays the
  reference's tests write (``_positions_arr``, ``_goals_arr``, ``_reached_arr`` ...): they are host
  mirrors of the device state -- uploaded at the entry of every ``reset``/``step``/``get_obs`` call,
  downloaded at exit.

The transition itself (moves, conflicts, observations, masks, lock metrics, blocking, rewards,
termination) always runs in ``libmapf_b200.so`` on the GPU: this class is a B = 1 view of the
batched kernels driven through the C ABI's host-buffer entry points.  No CUDA device -> it raises.

RNG: with ``rng_backend="numpy"`` (default, seed-compatible with the reference) the *draws* --
``rng.choice`` for reset layouts (ENV:277) and ``rng.integers`` inside ``_assign_new_goal``
(ENV:300) -- come from a host ``numpy.random.Generator`` exactly like the reference's and are fed
to the kernels through the replay hooks (``starts/goals`` overrides, ``goal_override``); with
``rng_backend="philox"`` they are drawn inside the kernels (Philox4x32-10).
"""
from __future__ import annotations

import ctypes as C
import logging

import numpy as np

from . import _native as nat
from . import maps
from .actions import DOWN, LEFT, NO_OP, RIGHT, UP
from .spaces import Box, Discrete, MultiAgentEnv, MultiBinary

logger = logging.getLogger(__name__)

_INFO_ALL_KEYS = (
    ("goals_reached_step", nat.I_GOALS_REACHED_STEP), ("goals_reached_total", nat.I_GOALS_REACHED_TOTAL),
    ("blocking_count_step", nat.I_BLOCKING_COUNT_STEP), ("blocking_count_total", nat.I_BLOCKING_COUNT_TOTAL),
    ("deadlock_step", nat.I_DEADLOCK_STEP), ("livelock_step", nat.I_LIVELOCK_STEP),
    ("deadlock_event_step", nat.I_DEADLOCK_EVENT_STEP), ("livelock_event_step", nat.I_LIVELOCK_EVENT_STEP),
    ("deadlock_events_total", nat.I_DEADLOCK_EVENTS_TOTAL), ("livelock_events_total", nat.I_LIVELOCK_EVENTS_TOTAL),
    ("deadlock_steps_total", nat.I_DEADLOCK_STEPS_TOTAL), ("livelock_steps_total", nat.I_LIVELOCK_STEPS_TOTAL),
)


def _vp(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class ReferenceModel(MultiAgentEnv):
    """Multi-agent grid world with flat per-agent observations (GPU-backed drop-in)."""

    EMPTY_CELL = 0
    OBSTACLE_CELL = 1
    OTHER_AGENT_CELL = 2
    OWN_GOAL_CELL = 3
    OTHER_GOAL_CELL = 4
    TRAVERSABLE_LOCAL_VALUES = (EMPTY_CELL, OWN_GOAL_CELL, OTHER_GOAL_CELL)
    UNASSIGNED_OWNER = -1

    # ------------------------------------------------------------------ construction, ENV:34-193
    def __init__(self, env_config):
        super().__init__()
        cfg = env_config
        self.step_count = 0
        self.steps_per_episode = cfg.get("steps_per_episode", 100)
        self._num_agents = int(cfg.get("num_agents", 2))
        self.sensor_range = cfg.get("sensor_range", 1)
        self.deterministic = cfg.get("deterministic", False)
        self.normalize_goal_delta = cfg.get("normalize_goal_delta", True)
        self.include_goal_distance = cfg.get("include_goal_distance", False)
        self.include_action_mask_in_obs = bool(cfg.get("include_action_mask_in_obs", False))
        self.include_blocking_pressure_in_obs = bool(cfg.get("include_blocking_pressure_in_obs", True))
        self.validate_observation_space = bool(cfg.get("validate_observation_space", False))
        self.possible_agents = [f"agent_{i}" for i in range(self._num_agents)]
        self.agents = self.possible_agents.copy()
        self.render_env = cfg.get("render_env", False)
        self.info_mode = str(cfg.get("info_mode", "lite")).lower()
        self.lifelong_mapf = bool(cfg.get("lifelong_mapf", False))
        self._needs_action_mask = self.include_action_mask_in_obs or self.info_mode == "full"
        if self.info_mode not in {"lite", "full"}:
            msg = f"Unsupported info_mode '{self.info_mode}'. Expected 'lite' or 'full'."
            raise ValueError(msg)
        self.deadlock_window_steps = max(1, int(cfg.get("deadlock_window_steps", 8)))
        self.livelock_window_steps = max(1, int(cfg.get("livelock_window_steps", 16)))
        self.lock_nearby_manhattan = max(1, int(cfg.get("lock_nearby_manhattan", 2)))
        self.lock_progress_epsilon = float(cfg.get("lock_progress_epsilon", 1))
        self.lock_min_neighbors = max(1, int(cfg.get("lock_min_neighbors", 1)))
        self.enable_lock_metrics = bool(cfg.get("enable_lock_metrics", True))
        self.goal_reached_once = dict.fromkeys(self.agents, False)
        self._episode_blocking_count = 0.0
        self._episode_deadlock_events = 0.0
        self._episode_livelock_events = 0.0
        self._episode_deadlock_steps = 0.0
        self._episode_livelock_steps = 0.0
        self._deadlock_state_prev = False
        self._livelock_state_prev = False
        self._episode_goals_reached_total = 0.0
        self._agent_index = {aid: i for i, aid in enumerate(self.agents)}
        self._coord_dtype = np.int16
        self._rng_backend = str(cfg.get("rng_backend", "numpy")).lower()
        if self._rng_backend not in {"numpy", "philox"}:
            raise ValueError(f"Unsupported rng_backend '{self._rng_backend}'. Expected 'numpy' or 'philox'.")

        self.seed = cfg.get("seed", None)
        self.rng = np.random.default_rng(self.seed) if self.seed is not None else np.random.default_rng()

        if cfg.get("grid") is not None:
            self.grid = np.ascontiguousarray(np.asarray(cfg["grid"]), dtype=np.uint8)
        else:
            self.grid = maps.get_grid(cfg["env_name"])
        if self.grid.ndim != 2:
            raise ValueError("ReferenceModel takes one [R, C] map")
        R, Cc = self.grid.shape
        N = self._num_agents
        self._free_positions = np.argwhere(self.grid == self.EMPTY_CELL).astype(self._coord_dtype, copy=False)

        # ---- device handle (B = 1)
        self._lib = nat.lib()
        dev = cfg.get("device", "cuda:0")
        dev_index = int(str(dev).split(":")[1]) if ":" in str(dev) else 0
        ccfg = nat.make_config(cfg, R, Cc, 1, dev_index, int(cfg.get("env_id_base", 0)), False)
        if ccfg.step_kernel == 0 and Cc <= 32 and R <= 64:
            # B = 1: launch-latency bound either way.  The env-per-thread kernel keeps the reference's owner-grid
            # semantics even for injected states with several agents on one cell (its occupancy board IS
            # "_occupancy_owner != -1", and agent_flags carries who owns a shared cell), so the wrapper -- the
            # place where tests inject such states -- takes it whenever the map fits (all eight reference maps do).
            ccfg.step_kernel = 2
        # Layouts reach the kernels as `starts`/`goals` overrides (deterministic table, numpy draw), so
        # the handle runs its restore-starts reset path; only the Philox backend draws in-kernel.
        in_kernel_draw = (not self.deterministic) and self._rng_backend == "philox"
        ccfg.deterministic = 0 if in_kernel_draw else 1
        if in_kernel_draw and self._free_positions.shape[0] < 2 * N:
            msg = (f"Environment has only {self._free_positions.shape[0]} free cells, "
                   f"but {2 * N} are required for starts and goals.")
            raise ValueError(msg)
        self._view_side = int(self.sensor_range) * 2 + 1
        h = C.c_void_p()
        nat.check(self._lib.mapf_create(C.byref(ccfg), C.byref(h)))
        self._h = h
        nat.check(self._lib.mapf_set_map(self._h, _vp(self.grid)))
        nat.check(self._lib.mapf_alloc_state(self._h))

        # ---- host images of the device state (mapf_state layout, B = 1) and their mirrors
        LW = self.livelock_window_steps
        self._hs = {
            "positions": np.zeros((1, N, 2), np.int16), "goals": np.zeros((1, N, 2), np.int16),
            "starts": np.zeros((1, N, 2), np.int16), "agent_flags": np.zeros((1, N), np.uint8),
            "lock_goal_progress": np.zeros((1, N), np.uint32), "lock_moved": np.zeros((1, N), np.uint32),
            "lock_failed_move": np.zeros((1, N), np.uint32), "lock_distance": np.zeros((1, LW, N), np.int16),
            "env_words": np.zeros((1, nat.ENV_WORDS), np.int32),
            "env_metrics": np.zeros((1, nat.METRIC_COUNT), np.float64),
        }
        self._starts_arr = self._hs["starts"][0]
        self._positions_arr = self._hs["positions"][0]
        self._goals_arr = self._hs["goals"][0]
        self._reached_arr = np.zeros(N, dtype=np.bool_)
        self._completed_once_arr = np.zeros(N, dtype=np.bool_)
        self._blocking_pressure_prev_arr = np.zeros(N, dtype=np.float32)
        self._occupancy_owner = np.full(self.grid.shape, self.UNASSIGNED_OWNER, dtype=np.int16)
        self._goal_owner = np.full(self.grid.shape, self.UNASSIGNED_OWNER, dtype=np.int16)
        self._scratch_intended_next = np.zeros((N, 2), np.int16)
        self._scratch_moved_flags = np.zeros(N, np.bool_)
        self._scratch_failed_move_flags = np.zeros(N, np.bool_)
        self._action_deltas = np.array([[0, 0], [-1, 0], [0, 1], [1, 0], [0, -1]], dtype=self._coord_dtype)
        self._lock_history_size = max(self.deadlock_window_steps, self.livelock_window_steps)
        self._lock_hist_count = 0
        V = self._view_side
        self._ho = {
            "local_obs": np.zeros((1, N, V, V), np.uint8), "action_mask": np.zeros((1, N, 5), np.int8),
            "goal_delta": np.zeros((1, N, 2), np.float32), "blocking_prev": np.zeros((1, N), np.uint8),
            "reward": np.zeros((1, N), np.float32), "terminated": np.zeros(1, np.uint8),
            "truncated": np.zeros(1, np.uint8), "step_flags": np.zeros(1, np.uint8),
            "agent_step_flags": np.zeros((1, N), np.uint8), "info": np.zeros((1, nat.INFO_WORDS), np.int32),
        }
        self._cout = nat.MapfOutputs(**{k: self._ho[k].ctypes.data for k in nat.OUTPUT_FIELDS})
        self._bind_public_state_views()

        if self.deterministic:  # ENV:124-132
            name = cfg.get("env_name")
            if cfg.get("starts") is not None and cfg.get("goals") is not None:
                starts = np.asarray(cfg["starts"], np.int16)
                goals = np.asarray(cfg["goals"], np.int16)
            else:
                sp, gp = maps.get_start_positions(name, N), maps.get_goal_positions(name, N)
                starts = np.array([sp[a] for a in self.agents], np.int16)
                goals = np.array([gp[a] for a in self.agents], np.int16)
            np.copyto(self._starts_arr, starts)
            np.copyto(self._goals_arr, goals)
            np.copyto(self._positions_arr, self._starts_arr)
            self._rebuild_goal_owner()
            self._rebuild_occupancy_owner()
            self._upload()
        else:
            self.generate_starts_goals()

        # ---- spaces, ENV:136-184
        self._local_obs_space = Box(low=0, high=self.OTHER_GOAL_CELL, shape=(V, V), dtype=np.uint8)
        gd_low = np.array([-(R - 1), -(Cc - 1)], dtype=np.float32)
        gd_high = np.array([R - 1, Cc - 1], dtype=np.float32)
        self._goal_delta_denominator = np.array([max(R - 1, 1), max(Cc - 1, 1)], dtype=np.float32)
        if self.normalize_goal_delta:
            gd_low = gd_low / self._goal_delta_denominator
            gd_high = gd_high / self._goal_delta_denominator
        self._goal_delta_space = Box(low=np.asarray(gd_low, np.float32), high=np.asarray(gd_high, np.float32),
                                     shape=(2,), dtype=np.float32)
        self._single_act_space = Discrete(5)
        self._action_mask_space = MultiBinary(int(self._single_act_space.n))
        self._blocking_pressure_space = Box(low=np.zeros(1, np.float32), high=np.ones(1, np.float32),
                                            dtype=np.float32)
        self._single_obs_space, self._obs_slices = self._build_obs_layout()
        self._single_obs_len = int(np.prod(self._single_obs_space.shape))
        self.observation_spaces = dict.fromkeys(self.possible_agents, self._single_obs_space)
        self.action_spaces = dict.fromkeys(self.possible_agents, self._single_act_space)
        self.observation_space = self._single_obs_space
        self.action_space = self._single_act_space

    # ------------------------------------------------------------------ small reference-compatible helpers
    def _bind_public_state_views(self):
        self.starts = {a: self._starts_arr[i] for a, i in self._agent_index.items()}
        self.positions = {a: self._positions_arr[i] for a, i in self._agent_index.items()}
        self.goals = {a: self._goals_arr[i] for a, i in self._agent_index.items()}

    def _rebuild_occupancy_owner(self):
        """Derived view (ENV:200-205): the kernels never store it, tests and hooks read it."""
        self._occupancy_owner.fill(self.UNASSIGNED_OWNER)
        p = self._positions_arr
        for idx in range(self._num_agents):
            self._occupancy_owner[p[idx, 0], p[idx, 1]] = idx

    def _rebuild_goal_owner(self):
        self._goal_owner.fill(self.UNASSIGNED_OWNER)
        g = self._goals_arr
        for idx in range(self._num_agents):
            self._goal_owner[g[idx, 0], g[idx, 1]] = idx

    def _build_obs_component_spaces(self):
        comps = [("local_obs", self._local_obs_space), ("goal_delta", self._goal_delta_space)]
        if self.include_goal_distance:
            top = float(np.abs(self._goal_delta_space.high).sum())
            comps.append(("goal_distance", Box(low=np.zeros(1, np.float32),
                                               high=np.asarray([top], np.float32), dtype=np.float32)))
        if self.include_blocking_pressure_in_obs:
            comps.append(("blocking_pressure_prev", self._blocking_pressure_space))
        if self.include_action_mask_in_obs:
            comps.append(("action_mask", self._action_mask_space))
        return comps

    def _build_obs_layout(self):
        lows, highs, slices, at = [], [], {}, 0
        for name, space in self._build_obs_component_spaces():
            if isinstance(space, MultiBinary):
                n = int(np.prod(space.shape))
                lo, hi = np.zeros(n, np.float32), np.ones(n, np.float32)
            else:
                lo = space.low.astype(np.float32).reshape(-1)
                hi = space.high.astype(np.float32).reshape(-1)
            slices[name] = slice(at, at + lo.size)
            at += lo.size
            lows.append(lo)
            highs.append(hi)
        space = Box(low=np.asarray(np.concatenate(lows), np.float32),
                    high=np.asarray(np.concatenate(highs), np.float32), dtype=np.float32)
        return space, slices

    # ------------------------------------------------------------------ host <-> device state
    def _pack_mirrors(self):
        af = (self._reached_arr.astype(np.uint8) * nat.AF_REACHED
              | self._completed_once_arr.astype(np.uint8) * nat.AF_COMPLETED_ONCE
              | (self._blocking_pressure_prev_arr != 0).astype(np.uint8) * nat.AF_BLOCKING_PREV)
        # the occupancy-owner grid is an input like the other mirrors (the reference's tests rebuild it after
        # writing positions, tests/test_reference_model_multi_agent_invariants.py:28-38): an agent that does not own
        # the cell it stands on (only possible when several were injected onto one cell) is flagged for the kernel
        p = self._positions_arr
        own = self._occupancy_owner[p[:, 0].astype(np.intp), p[:, 1].astype(np.intp)]
        af = af | ((own != np.arange(self._num_agents)).astype(np.uint8) * nat.AF_NOT_OWNER)
        self._hs["agent_flags"][0] = af
        w = self._hs["env_words"][0]
        w[nat.W_STEP_COUNT] = int(self.step_count)
        w[nat.W_GOALS_TOTAL] = int(self._episode_goals_reached_total)
        w[nat.W_BLOCKING_TOTAL] = int(self._episode_blocking_count)
        w[nat.W_DEADLOCK_EVENTS] = int(self._episode_deadlock_events)
        w[nat.W_LIVELOCK_EVENTS] = int(self._episode_livelock_events)
        w[nat.W_DEADLOCK_STEPS] = int(self._episode_deadlock_steps)
        w[nat.W_LIVELOCK_STEPS] = int(self._episode_livelock_steps)
        w[nat.W_LOCK_PREV] = int(bool(self._deadlock_state_prev)) | (int(bool(self._livelock_state_prev)) << 1)

    def _unpack_mirrors(self):
        af = self._hs["agent_flags"][0]
        self._reached_arr[:] = (af & nat.AF_REACHED) != 0
        self._completed_once_arr[:] = (af & nat.AF_COMPLETED_ONCE) != 0
        self._blocking_pressure_prev_arr[:] = ((af & nat.AF_BLOCKING_PREV) != 0).astype(np.float32)
        w = self._hs["env_words"][0]
        self.step_count = int(w[nat.W_STEP_COUNT])
        self._episode_goals_reached_total = float(w[nat.W_GOALS_TOTAL])
        self._episode_blocking_count = float(w[nat.W_BLOCKING_TOTAL])
        self._episode_deadlock_events = float(w[nat.W_DEADLOCK_EVENTS])
        self._episode_livelock_events = float(w[nat.W_LIVELOCK_EVENTS])
        self._episode_deadlock_steps = float(w[nat.W_DEADLOCK_STEPS])
        self._episode_livelock_steps = float(w[nat.W_LIVELOCK_STEPS])
        self._deadlock_state_prev = bool(w[nat.W_LOCK_PREV] & 1)
        self._livelock_state_prev = bool(w[nat.W_LOCK_PREV] & 2)
        self._lock_hist_count = min(int(w[nat.W_LOCK_COUNT]), self._lock_history_size)
        for a, i in self._agent_index.items():
            self.goal_reached_once[a] = bool(self._completed_once_arr[i])
        self._rebuild_goal_owner()
        self._occupancy_owner.fill(self.UNASSIGNED_OWNER)   # ENV:102: the grid outlives the step, non-owners included
        p = self._positions_arr
        for idx in range(self._num_agents):
            if not (af[idx] & nat.AF_NOT_OWNER):
                self._occupancy_owner[p[idx, 0], p[idx, 1]] = idx

    _MIRRORED = ("positions", "goals", "starts", "agent_flags", "env_words")

    def _state_struct(self, keys) -> nat.MapfState:
        return nat.MapfState(**{k: self._hs[k].ctypes.data for k in keys})

    def _upload(self, keys=None):
        self._pack_mirrors()
        st = self._state_struct(keys or self._MIRRORED)
        nat.check(self._lib.mapf_set_state_host(self._h, C.byref(st)))

    def _download(self, keys=None):
        st = self._state_struct(keys or self._MIRRORED)
        nat.check(self._lib.mapf_get_state_host(self._h, C.byref(st)))
        self._unpack_mirrors()

    def _reset_lock_tracking(self):
        """ENV:360-372: clear the lock history (device shift registers / ring) and its counters."""
        for k in ("lock_goal_progress", "lock_moved", "lock_failed_move", "lock_distance"):
            self._hs[k].fill(0)
        self._episode_deadlock_events = self._episode_livelock_events = 0.0
        self._episode_deadlock_steps = self._episode_livelock_steps = 0.0
        self._deadlock_state_prev = self._livelock_state_prev = False
        self._hs["env_words"][0, nat.W_LOCK_COUNT] = 0
        self._hs["env_words"][0, nat.W_LOCK_HEAD] = 0
        self._lock_hist_count = 0
        self._upload(self._MIRRORED + ("lock_goal_progress", "lock_moved", "lock_failed_move", "lock_distance"))

    # ------------------------------------------------------------------ RNG hooks (host draws, see module doc)
    def generate_starts_goals(self):
        """ENV:267-282: 2N distinct free cells, first N starts, next N goals."""
        need = self._num_agents * 2
        if self._free_positions.shape[0] < need:
            msg = (f"Environment has only {self._free_positions.shape[0]} free cells, "
                   f"but {need} are required for starts and goals.")
            raise ValueError(msg)
        if self._rng_backend == "philox":
            nat.check(self._lib.mapf_reset_host(self._h, None, None, None, None))
            self._download(self._MIRRORED)
            return
        picks = self.rng.choice(self._free_positions.shape[0], size=need, replace=False)
        np.copyto(self._starts_arr, self._free_positions[picks[: self._num_agents]])
        np.copyto(self._positions_arr, self._starts_arr)
        np.copyto(self._goals_arr, self._free_positions[picks[self._num_agents:]])
        self._rebuild_goal_owner()
        self._rebuild_occupancy_owner()
        self._upload()

    def _assign_new_goal(self, agent_idx: int) -> np.ndarray:
        """Host-side goal *draw* (ENV:284-304) for the numpy RNG backend / test monkeypatches.

        Called with the mirrors showing the mid-step state the reference's method would see
        (agents <= agent_idx moved, later ones not).  It only picks the cell; the transition that
        consumes it runs in the step kernel through ``goal_override``."""
        old = self._goals_arr[agent_idx]
        self._goal_owner[int(old[0]), int(old[1])] = self.UNASSIGNED_OWNER
        fy, fx = self._free_positions[:, 0], self._free_positions[:, 1]
        open_cells = (self._occupancy_owner[fy, fx] == self.UNASSIGNED_OWNER) & \
                     (self._goal_owner[fy, fx] == self.UNASSIGNED_OWNER)
        cand = np.flatnonzero(open_cells)
        if cand.size == 0:
            msg = "No valid cell available for lifelong goal reassignment."
            raise RuntimeError(msg)
        pick = int(cand[int(self.rng.integers(cand.size))])
        new_goal = self._free_positions[pick]
        self._goals_arr[agent_idx, :] = new_goal
        self._goal_owner[int(new_goal[0]), int(new_goal[1])] = agent_idx
        return new_goal

    # ------------------------------------------------------------------ observation packing
    def _get_goal_delta(self, agent_id: str) -> np.ndarray:
        self._observe()
        return self._ho["goal_delta"][0, self._agent_index[agent_id]].copy()

    def _flat_from_channels(self, idx: int) -> np.ndarray:
        """ENV:306-328 component order; channel values come from the kernels."""
        flat = np.empty(self._single_obs_len, dtype=np.float32)
        s = self._obs_slices
        flat[s["local_obs"]] = self._ho["local_obs"][0, idx].reshape(-1)
        gd = self._ho["goal_delta"][0, idx]
        flat[s["goal_delta"]] = gd
        if "goal_distance" in s:
            flat[s["goal_distance"]] = np.float32(np.abs(gd).sum(dtype=np.float32))
        if "blocking_pressure_prev" in s:
            flat[s["blocking_pressure_prev"]] = np.float32(self._ho["blocking_prev"][0, idx])
        if "action_mask" in s:
            flat[s["action_mask"]] = self._ho["action_mask"][0, idx]
        return flat

    def _flatten_observation(self, agent_id: str, local_obs=None, action_mask=None):
        idx = self._agent_index[agent_id]
        if local_obs is None:
            self._observe()
            return self._flat_from_channels(idx)
        flat = self._flat_from_channels(idx)
        flat[self._obs_slices["local_obs"]] = np.asarray(local_obs, np.float32).reshape(-1)
        if "action_mask" in self._obs_slices and action_mask is not None:
            flat[self._obs_slices["action_mask"]] = np.asarray(action_mask, np.float32).reshape(-1)
        return flat

    def _coerce_and_validate_observation(self, agent_id: str, obs: np.ndarray, *, where: str) -> np.ndarray:
        obs = np.asarray(obs, dtype=np.float32)
        if self.validate_observation_space and not self.observation_space.contains(obs):
            msg = (f"{where} produced invalid observation for {agent_id} "
                   f"(dtype={obs.dtype}, min={float(np.min(obs))}, max={float(np.max(obs))}).")
            raise ValueError(msg)
        return obs

    def _build_full_info(self, agent_id: str, idx: int) -> dict:
        return {  # ENV:350-358
            "position": np.asarray(self.positions[agent_id]),
            "goal": np.asarray(self.goals[agent_id]),
            "goal_delta": self._ho["goal_delta"][0, idx].copy(),
            "action_mask": self._ho["action_mask"][0, idx].copy(),
            "local_obs": self._ho["local_obs"][0, idx].copy(),
        }

    def _collect_obs(self, where: str, info: dict):
        obs = {}
        for idx, aid in enumerate(self.agents):
            obs[aid] = self._coerce_and_validate_observation(aid, self._flat_from_channels(idx), where=where)
            if self.info_mode == "full":
                info[aid] = self._build_full_info(aid, idx)
        return obs

    def _observe(self):
        self._upload()
        nat.check(self._lib.mapf_observe_host(self._h, C.byref(self._cout)))

    # ------------------------------------------------------------------ reset, ENV:440-472
    def reset(self, *, seed=None, options=None):
        infos = {aid: {} for aid in self.agents}
        self.goal_reached_once = dict.fromkeys(self.agents, False)
        if self.deterministic:
            self._upload()  # tests may have edited `_starts_arr` / `_goals_arr`
            nat.check(self._lib.mapf_reset_host(self._h, None, None, None, C.byref(self._cout)))
        elif self._rng_backend == "philox":
            need = self._num_agents * 2
            if self._free_positions.shape[0] < need:
                msg = (f"Environment has only {self._free_positions.shape[0]} free cells, "
                       f"but {need} are required for starts and goals.")
                raise ValueError(msg)
            nat.check(self._lib.mapf_reset_host(self._h, None, None, None, C.byref(self._cout)))
        else:
            self.generate_starts_goals()
            nat.check(self._lib.mapf_reset_host(self._h, None, _vp(self._hs["starts"]), _vp(self._hs["goals"]),
                                                C.byref(self._cout)))
        self._download()
        obs = self._collect_obs("reset", infos)
        if self.render_env:
            self.render()
        return obs, infos

    # ------------------------------------------------------------------ step, ENV:474-695
    def _goal_hook_active(self) -> bool:
        return self.lifelong_mapf and (self._rng_backend == "numpy" or "_assign_new_goal" in self.__dict__)

    def step(self, action_dict):
        N = self._num_agents
        if not action_dict or any(aid not in action_dict for aid in self.agents):
            action_dict = dict.fromkeys(self.agents, NO_OP)
            logger.warning("No actions provided or missing agent actions. Defaulting to no-op actions: %s", action_dict)
        acts = np.zeros((1, N), np.int8)
        for idx, aid in enumerate(self.agents):
            action = int(action_dict[aid])
            if action < NO_OP or action > LEFT:
                self.step_count += 1  # the reference has already counted the step when it raises (ENV:475)
                msg = f"Invalid action {action} for {aid}"
                raise ValueError(msg)
            acts[0, idx] = action
        prev = self._positions_arr.copy()
        self._upload()
        override = None
        if self._goal_hook_active():
            override = self._draw_goal_overrides(acts, prev)
        nat.check(self._lib.mapf_step_host(self._h, _vp(acts), _vp(override), None, C.byref(self._cout), 0))
        self._download()
        bits = C.c_uint32(0)
        nat.check(self._lib.mapf_poll_errors(self._h, C.byref(bits), None))
        if bits.value & nat.DEV_ERR_NO_GOAL_CELL:
            msg = "No valid cell available for lifelong goal reassignment."
            raise RuntimeError(msg)
        return self._results(acts[0], prev)

    def _draw_goal_overrides(self, acts: np.ndarray, prev: np.ndarray):
        """Who arrives this step is decided by the moves alone (ENV:502-563: they depend on the owner grid, not on
        anybody's goal), so the wrapper replays the move loop on its host mirrors -- no probe launch, no roll-back --
        lets the host RNG hook pick the new cells in the reference's order and hands them to the one real launch as
        overrides."""
        N = self._num_agents
        R, Cc = self.grid.shape
        mirrors = (self._reached_arr.copy(), self._completed_once_arr.copy(), self._blocking_pressure_prev_arr.copy(),
                   self._episode_goals_reached_total)
        saved = {"positions": self._positions_arr.copy(), "goals": self._goals_arr.copy(),
                 "owner": self._occupancy_owner.copy()}
        owner = self._occupancy_owner.copy()
        new_pos = prev.copy()
        for i in range(N):   # ENV:512-526
            a = int(acts[0, i])
            if a == NO_OP:
                continue
            tr, tc = (int(v) for v in prev[i] + self._action_deltas[a])
            if not (0 <= tr < R and 0 <= tc < Cc) or self.grid[tr, tc] != 0 or owner[tr, tc] != self.UNASSIGNED_OWNER:
                continue
            owner[int(prev[i, 0]), int(prev[i, 1])] = self.UNASSIGNED_OWNER
            owner[tr, tc] = i
            new_pos[i] = (tr, tc)
        arrived = (new_pos == self._goals_arr).all(axis=1)
        if not arrived.any():
            return None
        override = np.full((1, N, 2), -1, np.int16)
        order = np.arange(N)
        for idx in np.flatnonzero(arrived):
            self._positions_arr[:] = np.where((order <= idx)[:, None], new_pos, prev)  # snapshot idx
            self._rebuild_occupancy_owner()
            self._rebuild_goal_owner()
            self._completed_once_arr[idx] = True       # ENV:550-553, visible to a patched hook
            self._reached_arr[idx] = False
            self._episode_goals_reached_total += 1.0
            self.goal_reached_once[self.agents[idx]] = True
            self._assign_new_goal(int(idx))
            override[0, idx] = self._goals_arr[idx]
        self._positions_arr[:] = saved["positions"]
        self._goals_arr[:] = saved["goals"]
        self._occupancy_owner[:] = saved["owner"]
        self._rebuild_goal_owner()
        self._reached_arr[:], self._completed_once_arr[:], self._blocking_pressure_prev_arr[:] = mirrors[:3]
        self._episode_goals_reached_total = mirrors[3]
        return override

    def _results(self, acts: np.ndarray, prev: np.ndarray):
        N = self._num_agents
        asf = self._ho["agent_step_flags"][0]
        iw = self._ho["info"][0]
        self._scratch_intended_next[:] = prev + self._action_deltas[acts.astype(np.int64)]
        self._scratch_moved_flags[:] = (asf & nat.ASF_MOVED) != 0
        self._scratch_failed_move_flags[:] = (asf & nat.ASF_FAILED_MOVE) != 0
        info = {aid: {} for aid in self.agents}
        obs = self._collect_obs("step", info)
        rewards = {aid: float(self._ho["reward"][0, i]) for i, aid in enumerate(self.agents)}
        goals_total = float(iw[nat.I_GOALS_REACHED_TOTAL])
        blocking_total = float(iw[nat.I_BLOCKING_COUNT_TOTAL])
        for i, aid in enumerate(self.agents):
            d = info[aid]
            d["blocking"] = float((asf[i] & nat.ASF_BLOCKING) != 0)
            d["goal_reached_step"] = float((asf[i] & nat.ASF_GOAL_REACHED) != 0)
            d["goals_reached_total"] = goals_total
            d["blocking_count_total"] = blocking_total
        info_all = {k: float(iw[j]) for k, j in _INFO_ALL_KEYS}
        if self.lifelong_mapf:  # ENV:638,653-655: float64 ratios formed on the host like the reference
            info_all["completion_ratio"] = float(np.mean(self._completed_once_arr))
            info_all["throughput"] = goals_total / float(max(self.step_count, 1))
        info["__all__"] = info_all
        term, trunc = bool(self._ho["terminated"][0]), bool(self._ho["truncated"][0])
        terminated = dict.fromkeys(self.agents, term)
        truncated = dict.fromkeys(self.agents, trunc)
        terminated["__all__"] = term
        truncated["__all__"] = trunc
        # ENV:659-666: co-located agents are reported like the reference does
        if N > 1:
            lin = self._positions_arr[:, 0].astype(np.int64) * self.grid.shape[1] + self._positions_arr[:, 1]
            if np.unique(lin).size != N:
                logger.warning("Agents occupy the same position: %s", self._positions_arr.tolist())
        if self.render_env:
            self.render()
        return obs, rewards, terminated, truncated, info

    # ------------------------------------------------------------------ public getters, ENV:697-773
    def get_next_position(self, action: int, pos):
        action = int(action)
        if action < NO_OP or action > LEFT:
            msg = "Invalid action"
            raise ValueError(msg)
        return np.asarray(pos, dtype=self._coord_dtype) + self._action_deltas[action]

    def get_obs(self, agent_id: str):
        """Local observation of one agent from the CURRENT state (kernel: mapf_observe)."""
        self._observe()
        return self._ho["local_obs"][0, self._agent_index[agent_id]].copy()

    def get_action_mask(self, obs):
        """Mask implied by a given local observation array (ENV:749-773): a pure function of its
        argument -- the centre's four neighbours are traversable iff their code is 0, 3 or 4."""
        obs = np.asarray(obs)
        c = int(self.sensor_range)
        mask = np.zeros(self._action_mask_space.shape, dtype=self._action_mask_space.dtype)
        mask[NO_OP] = 1
        ok = self.TRAVERSABLE_LOCAL_VALUES
        if c > 0 and obs[c - 1, c] in ok:
            mask[UP] = 1
        if c < obs.shape[1] - 1 and obs[c, c + 1] in ok:
            mask[RIGHT] = 1
        if c < obs.shape[0] - 1 and obs[c + 1, c] in ok:
            mask[DOWN] = 1
        if c > 0 and obs[c, c - 1] in ok:
            mask[LEFT] = 1
        return mask

    def render(self, mode="human"):
        """Text rendering (the reference's matplotlib view, ENV:775-916, is host-side debug code)."""
        chars = np.where(self.grid == 1, "#", ".").astype(object)
        for i in range(self._num_agents):
            gy, gx = self._goals_arr[i]
            chars[gy, gx] = chr(ord("a") + i % 26)
        for i in range(self._num_agents):
            py, px = self._positions_arr[i]
            chars[py, px] = chr(ord("A") + i % 26)
        text = "\n".join("".join(row) for row in chars)
        if mode == "human":
            print(text)
        return text

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.mapf_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
