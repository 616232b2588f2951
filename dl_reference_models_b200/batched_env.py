"""Batched tensor API of the MAPF environment: B independent envs of one GPU stepped by one
kernel launch (``libmapf_b200.so``, sm_100a).

This is the on-device rollout interface added beside the reference's RLlib ``MultiAgentEnv``
API (see :mod:`dl_reference_models_b200.reference_model` for the drop-in wrapper).  State and
outputs are ``torch`` CUDA tensors owned by this object; torch is only the allocator / stream
provider, the transition itself runs in the hand-written kernels.

Semantics follow ``src/environments/reference_model_multi_agent.py`` of the reference
(``ENV:line``): sequential agent-index move resolution (ENV:502-526), staggered observations
(ENV:528-536), lifelong goal reassignment (ENV:284-304, 547-556), lock metrics (ENV:389-438,
577-606), blocking (ENV:608-625), rewards and termination (ENV:658-690).
"""
from __future__ import annotations

import ctypes as C
from typing import NamedTuple

import numpy as np
import torch

from . import _native as nat
from . import maps


class StepOutput(NamedTuple):
    """Views of the env's output tensors (overwritten by the next reset/step call)."""

    local_obs: torch.Tensor         # uint8  [B,N,V,V]  ENV:707-747
    goal_delta: torch.Tensor        # float32[B,N,2]    ENV:330-335
    blocking_prev: torch.Tensor     # uint8  [B,N]      ENV:322
    action_mask: torch.Tensor       # int8   [B,N,5]    ENV:749-773
    reward: torch.Tensor            # float32[B,N]
    terminated: torch.Tensor        # uint8  [B]
    truncated: torch.Tensor         # uint8  [B]
    step_flags: torch.Tensor        # uint8  [B]   nat.SF_*
    agent_step_flags: torch.Tensor  # uint8  [B,N] nat.ASF_*
    info: torch.Tensor              # int32  [B,16] nat.I_*


def resolve_grid(env_config: dict) -> np.ndarray:
    """``grid`` (inline uint8 map, [R,C] or [B,R,C]) wins over ``env_name`` (ENV:80)."""
    if env_config.get("grid") is not None:
        return np.ascontiguousarray(np.asarray(env_config["grid"]), dtype=np.uint8)
    return maps.get_grid(env_config["env_name"])


class BatchedMapfEnv:
    def __init__(self, env_config: dict, num_envs: int, device="cuda:0", env_id_base: int = 0):
        self.env_config = dict(env_config)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("BatchedMapfEnv runs on a CUDA device only (no CPU fallback)")
        if not torch.cuda.is_available():
            raise RuntimeError("no CUDA device available: dl_reference_models_b200 has no CPU fallback")
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", dev_index)
        self._lib = nat.lib()
        grid = resolve_grid(self.env_config)
        self.per_env_maps = grid.ndim == 3
        if self.per_env_maps and grid.shape[0] != num_envs:
            raise ValueError(f"per-env grid has {grid.shape[0]} maps for {num_envs} envs")
        self.grid = grid
        R, Cc = grid.shape[-2:]
        self.cfg = nat.make_config(self.env_config, R, Cc, num_envs, dev_index, env_id_base, self.per_env_maps)
        self.B, self.N = int(num_envs), int(self.cfg.num_agents)
        self.V = 2 * int(self.cfg.sensor_range) + 1
        self.lifelong = bool(self.cfg.lifelong_mapf)
        self.deterministic = bool(self.cfg.deterministic)
        self.include_goal_distance = bool(self.env_config.get("include_goal_distance", False))
        self.include_blocking_pressure = bool(self.env_config.get("include_blocking_pressure_in_obs", True))
        self.include_action_mask = bool(self.env_config.get("include_action_mask_in_obs", False))

        h = C.c_void_p()
        nat.check(self._lib.mapf_create(C.byref(self.cfg), C.byref(h)))
        self._h = h
        nat.check(self._lib.mapf_set_map(self._h, grid.ctypes.data_as(C.c_void_p)))

        B, N, V, LW = self.B, self.N, self.V, int(self.cfg.livelock_window_steps)
        dev = self.device
        z = lambda shape, dt: torch.zeros(shape, dtype=dt, device=dev)  # noqa: E731
        self.state = {
            "positions": z((B, N, 2), torch.int16), "goals": z((B, N, 2), torch.int16),
            "starts": z((B, N, 2), torch.int16), "agent_flags": z((B, N), torch.uint8),
            "lock_goal_progress": z((B, N), torch.int32), "lock_moved": z((B, N), torch.int32),
            "lock_failed_move": z((B, N), torch.int32), "lock_distance": z((B, LW, N), torch.int16),
            "env_words": z((B, nat.ENV_WORDS), torch.int32),
            "env_metrics": z((B, nat.METRIC_COUNT), torch.float64),
        }
        sizes = (C.c_int64 * 10)()
        nat.check(self._lib.mapf_state_nbytes(self._h, C.byref(sizes)))
        for k, n in zip(nat.STATE_FIELDS, sizes):
            t = self.state[k]
            assert t.numel() * t.element_size() == n, (k, t.shape, n)
        self._cstate = nat.MapfState(**{k: self.state[k].data_ptr() for k in nat.STATE_FIELDS})
        nat.check(self._lib.mapf_bind_state(self._h, C.byref(self._cstate)))

        self.out = {
            "local_obs": z((B, N, V, V), torch.uint8), "action_mask": z((B, N, 5), torch.int8),
            "goal_delta": z((B, N, 2), torch.float32), "blocking_prev": z((B, N), torch.uint8),
            "reward": z((B, N), torch.float32), "terminated": z((B,), torch.uint8),
            "truncated": z((B,), torch.uint8), "step_flags": z((B,), torch.uint8),
            "agent_step_flags": z((B, N), torch.uint8), "info": z((B, nat.INFO_WORDS), torch.int32),
        }
        self._cout = nat.MapfOutputs(**{k: self.out[k].data_ptr() for k in nat.OUTPUT_FIELDS})
        self._flat = None
        self._metrics_dev = torch.zeros(nat.METRIC_COUNT, dtype=torch.float64, device=dev)
        self._actions = torch.zeros((B, N), dtype=torch.int8, device=dev)
        self._sample_counter = 0

        if self.deterministic:  # ENV:124-132: table layout installed once by the constructor
            starts = self.env_config.get("starts")
            goals = self.env_config.get("goals")
            if starts is None or goals is None:
                name = self.env_config["env_name"]
                sp = maps.get_start_positions(name, N)
                gp = maps.get_goal_positions(name, N)
                starts = np.array([sp[f"agent_{i}"] for i in range(N)], np.int16)
                goals = np.array([gp[f"agent_{i}"] for i in range(N)], np.int16)
            self.set_layout(starts, goals)

    # ------------------------------------------------------------------ lifecycle
    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.mapf_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _dev_tensor(self, x, dtype, shape):
        if x is None:
            return None
        t = torch.as_tensor(x, dtype=dtype) if not isinstance(x, torch.Tensor) else x.to(dtype)
        if tuple(t.shape) != tuple(shape):
            if t.ndim == len(shape) - 1 and tuple(t.shape) == tuple(shape[1:]):
                t = t.unsqueeze(0).expand(*shape)
            else:
                raise ValueError(f"expected shape {tuple(shape)}, got {tuple(t.shape)}")
        return t.to(self.device).contiguous()

    def _output(self) -> StepOutput:
        o = self.out
        return StepOutput(o["local_obs"], o["goal_delta"], o["blocking_prev"], o["action_mask"], o["reward"],
                          o["terminated"], o["truncated"], o["step_flags"], o["agent_step_flags"], o["info"])

    @staticmethod
    def _ptr(t):
        return None if t is None else C.c_void_p(t.data_ptr())

    # ------------------------------------------------------------------ transition
    def set_layout(self, starts, goals):
        """Install starts/goals (int16 [N,2] for all envs or [B,N,2]) and reset every env to them."""
        return self.reset(starts=starts, goals=goals)

    def reset(self, mask=None, starts=None, goals=None) -> StepOutput:
        """ENV:440-472 for the envs selected by ``mask`` (uint8/bool [B]; None = all).

        With ``starts``/``goals`` the layout is taken from them (what ``rng.choice`` produced in the
        reference, ENV:277-280); otherwise deterministic configs restore ``starts`` and keep the
        goals (ENV:452-455), the others draw 2N distinct free cells with Philox."""
        B, N = self.B, self.N
        m = self._dev_tensor(mask, torch.uint8, (B,))
        s = self._dev_tensor(starts, torch.int16, (B, N, 2))
        g = self._dev_tensor(goals, torch.int16, (B, N, 2))
        nat.check(self._lib.mapf_reset(self._h, self._ptr(m), self._ptr(s), self._ptr(g),
                                       C.byref(self._cout), self._stream()))
        return self._output()

    def step(self, actions=None, goal_override=None, goal_rank=None, auto_reset: bool = False, *, reward_out=None,
             terminated_out=None, truncated_out=None) -> StepOutput:
        """ENV:474-695 for every env.  ``actions``: int8 [B,N] in 0..4 (None = all NO_OP).

        ``goal_rank`` (int32 [B,N], >= 0 replaces ``rng.integers(n)`` of ENV:300) and
        ``goal_override`` (int16 [B,N,2], row >= 0 replaces the drawn cell) are the replay hooks
        that make lifelong runs bit-reproducible against a recorded reference trace.

        ``reward_out`` (float32 [B,N]), ``terminated_out`` / ``truncated_out`` (uint8 [B]): contiguous device tensors
        the kernel writes those channels to instead of the env's own buffers -- a rollout loop hands in the rows of
        its [T,...] buffers and needs no copy launches."""
        B, N = self.B, self.N
        a = self._dev_tensor(actions, torch.int8, (B, N))
        go = self._dev_tensor(goal_override, torch.int16, (B, N, 2))
        gr = self._dev_tensor(goal_rank, torch.int32, (B, N))
        # with the fused sampler the kernel may read this step's actions from the very buffer it refills
        # for the next step: every lane reads its own element before it writes it, so in-place is safe
        cout, over = self._cout, {}
        for key, t, dt, shape in (("reward", reward_out, torch.float32, (B, N)), ("terminated", terminated_out, torch.uint8, (B,)),
                                  ("truncated", truncated_out, torch.uint8, (B,))):
            if t is None:
                continue
            if t.dtype != dt or tuple(t.shape) != shape or not t.is_contiguous() or t.device != self.device:
                raise ValueError(f"{key}_out must be a contiguous {dt} tensor of shape {shape} on {self.device}")
            over[key] = t
        if over:
            cout = nat.MapfOutputs(**{k: (over[k] if k in over else self.out[k]).data_ptr() for k in nat.OUTPUT_FIELDS})
        nat.check(self._lib.mapf_step(self._h, self._ptr(a), self._ptr(go), self._ptr(gr),
                                      C.byref(cout), int(bool(auto_reset)), self._stream()))
        if getattr(self, "_fused", 0):
            self._sample_counter += 1
        if over:
            o = dict(self.out, **over)
            return StepOutput(o["local_obs"], o["goal_delta"], o["blocking_prev"], o["action_mask"], o["reward"],
                              o["terminated"], o["truncated"], o["step_flags"], o["agent_step_flags"], o["info"])
        return self._output()

    def step_many(self, steps: int, actions=None, auto_reset: bool = True, out: dict | None = None) -> dict:
        """``steps`` env steps with the fused sampler's actions (:meth:`fuse_sampler` first), in one kernel launch for
        small batches (``mapf_step_many``).  ``out``: optional dict of [steps, B, ...] device tensors keyed like
        ``self.out`` -- every step's outputs are kept (a rollout); without it only the last step's are, in the env's
        own buffers.  Returns the dict written to."""
        if not getattr(self, "_fused", 0):
            raise RuntimeError("step_many draws the actions of steps 2..K inside the launch: call fuse_sampler() first")
        a = self._actions if actions is None else self._dev_tensor(actions, torch.int8, (self.B, self.N))
        if out is None:
            cout, stride, ret = self._cout, 0, self.out
        else:
            for k in nat.OUTPUT_FIELDS:
                t = out[k]
                want = (int(steps),) + tuple(self.out[k].shape)
                if tuple(t.shape) != want or t.dtype != self.out[k].dtype or not t.is_contiguous() or t.device != self.device:
                    raise ValueError(f"out[{k!r}] must be a contiguous {self.out[k].dtype} tensor of shape {want} on {self.device}")
            cout, stride, ret = nat.MapfOutputs(**{k: out[k].data_ptr() for k in nat.OUTPUT_FIELDS}), self.B, out
        nat.check(self._lib.mapf_step_many(self._h, self._ptr(a), C.byref(cout), int(steps), stride, int(bool(auto_reset)),
                                           self._stream()))
        self._sample_counter += int(steps)
        return ret

    def rollout_buffers(self, steps: int) -> dict:
        """[steps, B, ...] device tensors for :meth:`step_many`."""
        return {k: torch.zeros((int(steps),) + tuple(v.shape), dtype=v.dtype, device=self.device) for k, v in self.out.items()}

    # ------------------------------------------------------------------ host-buffer transition
    HOST_CHANNELS = ("local_obs", "action_mask", "goal_delta", "blocking_prev", "reward", "terminated", "truncated")

    def host_buffers(self, channels=None) -> dict:
        """Pinned host arrays (numpy views of pinned torch tensors) for :meth:`step_host` / :meth:`reset_host`,
        one per requested output channel (default: what an RL worker consumes, ``HOST_CHANNELS``)."""
        channels = self.HOST_CHANNELS if channels is None else tuple(channels)
        unknown = set(channels) - set(nat.OUTPUT_FIELDS)
        if unknown:
            raise KeyError(f"unknown output channels {sorted(unknown)}")
        self._host_t = {k: torch.empty_like(self.out[k], device="cpu").pin_memory() for k in channels}
        self._host_actions = torch.zeros((self.B, self.N), dtype=torch.int8).pin_memory()
        self._host_np = {k: v.numpy() for k, v in self._host_t.items()}
        self._chost = nat.MapfOutputs(**{k: (self._host_t[k].data_ptr() if k in self._host_t else None)
                                         for k in nat.OUTPUT_FIELDS})
        return self._host_np

    def reset_host(self, mask=None, starts=None, goals=None) -> dict:
        """:meth:`reset` with host arrays in and out (``mapf_reset_host``): returns the dict of :meth:`host_buffers`."""
        if getattr(self, "_chost", None) is None:
            self.host_buffers()
        B, N = self.B, self.N
        m = None if mask is None else np.ascontiguousarray(np.asarray(mask, dtype=np.uint8).reshape(B))
        s = g = None
        if starts is not None or goals is not None:
            s = np.ascontiguousarray(np.broadcast_to(np.asarray(starts, dtype=np.int16), (B, N, 2)))
            g = np.ascontiguousarray(np.broadcast_to(np.asarray(goals, dtype=np.int16), (B, N, 2)))
        vp = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)  # noqa: E731
        nat.check(self._lib.mapf_host_wait_stream(self._h, self._stream()))   # torch work queued before this call
        nat.check(self._lib.mapf_reset_host(self._h, vp(m), vp(s), vp(g), C.byref(self._chost)))
        return self._host_np

    def step_host(self, actions=None, auto_reset: bool = True) -> dict:
        """One env step with HOST arrays in and out -- the vector-env call of a CPU-side RL worker
        (``mapf_step_host``: H2D of the actions, the step kernel, and the device-to-host delivery of the requested
        channels all happen inside the call; big batches cross PCIe bit-packed and are expanded into the returned
        arrays by the library's host threads).  ``actions``: integer array [B,N] in 0..4 (None = all
        NO_OP).  Returns the dict of :meth:`host_buffers` (the same arrays every call, overwritten in place)."""
        if getattr(self, "_chost", None) is None:
            self.host_buffers()
        a = None
        if actions is not None:
            src = actions.numpy() if isinstance(actions, torch.Tensor) else np.asarray(actions)
            if src.shape != (self.B, self.N):
                raise ValueError(f"expected actions of shape {(self.B, self.N)}, got {src.shape}")
            np.copyto(self._host_actions.numpy(), src, casting="unsafe")
            a = C.c_void_p(self._host_actions.data_ptr())
        nat.check(self._lib.mapf_host_wait_stream(self._h, self._stream()))   # torch work queued before this call
        nat.check(self._lib.mapf_step_host(self._h, a, None, None, C.byref(self._chost), int(bool(auto_reset))))
        return self._host_np

    def step_host_records(self, actions=None, auto_reset: bool = True) -> tuple:
        """Compact delivery (``mapf_step_host_records``): returns ``(records, small)`` -- ``records`` a pinned uint8 array
        holding the whole batch's bit-packed agent records (13 B per agent at sensor range 2; expand with
        :meth:`unpack_records` when and where the plain arrays are needed), ``small`` the dict of the remaining host
        channels.  Nothing is expanded inside the call."""
        if getattr(self, "_rec_t", None) is None:
            rs = int(self._lib.mapf_packed_record_bytes(self.V * self.V))
            self._rec_t = torch.empty((self.B * self.N * rs,), dtype=torch.uint8).pin_memory()
            small = ("blocking_prev", "terminated", "truncated", "step_flags", "agent_step_flags", "info")
            self._small_t = {k: torch.empty_like(self.out[k], device="cpu").pin_memory() for k in small}
            self._small_np = {k: v.numpy() for k, v in self._small_t.items()}
            self._csmall = nat.MapfOutputs(**{k: (self._small_t[k].data_ptr() if k in self._small_t else None)
                                              for k in nat.OUTPUT_FIELDS})
            self._rec_actions = torch.zeros((self.B, self.N), dtype=torch.int8).pin_memory()
        a = None
        if actions is not None:
            src = actions.numpy() if isinstance(actions, torch.Tensor) else np.asarray(actions)
            np.copyto(self._rec_actions.numpy(), src, casting="unsafe")
            a = C.c_void_p(self._rec_actions.data_ptr())
        nat.check(self._lib.mapf_host_wait_stream(self._h, self._stream()))
        nat.check(self._lib.mapf_step_host_records(self._h, a, C.c_void_p(self._rec_t.data_ptr()), C.byref(self._csmall),
                                                   int(bool(auto_reset))))
        return self._rec_t.numpy(), self._small_np

    def unpack_records(self, records: np.ndarray, threads: int = 4) -> dict:
        """Expand a records block of :meth:`step_host_records` into plain arrays (``mapf_unpack_records``, host only)."""
        B, N, V = self.B, self.N, self.V
        out = {"local_obs": np.empty((B, N, V, V), np.uint8), "action_mask": np.empty((B, N, 5), np.int8),
               "goal_delta": np.empty((B, N, 2), np.float32), "reward": np.empty((B, N), np.float32)}
        R, Cc = self.grid.shape[-2:]
        norm = bool(self.cfg.normalize_goal_delta)
        den0, den1 = (float(max(R - 1, 1)), float(max(Cc - 1, 1))) if norm else (1.0, 1.0)
        vp = lambda x: x.ctypes.data_as(C.c_void_p)  # noqa: E731
        nat.check(self._lib.mapf_unpack_records(vp(records), B * N, V * V, int(threads), vp(out["local_obs"]),
                                                vp(out["action_mask"]), vp(out["goal_delta"]), vp(out["reward"]),
                                                C.c_float(den0), C.c_float(den1)))
        return out

    def host_transfer_bytes(self) -> tuple:
        """(host-to-device, device-to-host) bytes that crossed PCIe in the last :meth:`step_host`."""
        h2d, d2h = C.c_int64(0), C.c_int64(0)
        nat.check(self._lib.mapf_host_transfer_bytes(self._h, C.byref(h2d), C.byref(d2h)))
        return int(h2d.value), int(d2h.value)

    def sample_actions(self, masked: bool = True) -> torch.Tensor:
        """Uniform (masked) random actions on device, the samplers of the reference's benchmark
        script (scripts/benchmark_multi_agent_env.py:38-57)."""
        self._sample_counter += 1
        if masked:
            nat.check(self._lib.mapf_sample_masked_actions(
                self._h, self._ptr(self.out["action_mask"]), self._ptr(self._actions),
                C.c_uint64(self._sample_counter), self._stream()))
        else:
            nat.check(self._lib.mapf_sample_random_actions(
                self._h, self._ptr(self._actions), C.c_uint64(self._sample_counter), self._stream()))
        return self._actions

    def fuse_sampler(self, mode: str | None = "masked") -> torch.Tensor:
        """Fold the benchmark's action sampler into the step launch: from now on every ``step``
        also writes the actions for the next step (uniform over the new mask, or over 0..4) into the
        returned int8 [B,N] tensor -- one launch per env step instead of two.  ``mode=None`` turns it
        off.  Draws are identical to :meth:`sample_actions` with the same call counter."""
        code = {None: 0, "masked": 1, "random": 2}[mode]
        nat.check(self._lib.mapf_set_fused_sampler(self._h, self._ptr(self._actions) if code else None, code,
                                                   C.c_uint64(self._sample_counter + 1)))
        self._fused = code
        return self._actions

    # ------------------------------------------------------------------ observations
    def flat_obs_dim(self, include_goal_distance=None, include_blocking_pressure=None, include_action_mask=None) -> int:
        gd, bp, am = self._flat_flags(include_goal_distance, include_blocking_pressure, include_action_mask)
        return int(self._lib.mapf_flat_obs_dim(self._h, gd, bp, am))

    def _flat_flags(self, gd, bp, am):
        gd = self.include_goal_distance if gd is None else gd
        bp = self.include_blocking_pressure if bp is None else bp
        am = self.include_action_mask if am is None else am
        return int(bool(gd)), int(bool(bp)), int(bool(am))

    def flat_obs(self, include_goal_distance=None, include_blocking_pressure=None, include_action_mask=None,
                 out: torch.Tensor | None = None) -> torch.Tensor:
        """float32 [B,N,D] in the reference's component order (ENV:214-236, 306-328)."""
        gd, bp, am = self._flat_flags(include_goal_distance, include_blocking_pressure, include_action_mask)
        D = int(self._lib.mapf_flat_obs_dim(self._h, gd, bp, am))
        if out is None:
            if self._flat is None or self._flat.shape[-1] != D:
                self._flat = torch.empty((self.B, self.N, D), dtype=torch.float32, device=self.device)
            out = self._flat
        nat.check(self._lib.mapf_pack_flat_obs(self._h, C.byref(self._cout), gd, bp, am,
                                               C.c_void_p(out.data_ptr()), self._stream()))
        return out

    # ------------------------------------------------------------------ metrics / errors / state
    def metrics_vector(self) -> torch.Tensor:
        """Episode-end metric sums of this shard: float64 [16] on device (see nat.METRIC_NAMES)."""
        nat.check(self._lib.mapf_metrics_reduce(self._h, self._ptr(self._metrics_dev), self._stream()))
        return self._metrics_dev

    def accumulate_occupancy(self, counts: torch.Tensor, active: torch.Tensor | None = None) -> torch.Tensor:
        """counts[r, c] (int64 [R,C] on this device) += agents standing on (r, c) over the envs with ``active`` != 0
        (uint8 [B], None = all): the occupancy heat-map of the reference's evaluator (main.py:153-155, 265-267)."""
        assert counts.dtype == torch.int64 and counts.is_contiguous() and tuple(counts.shape) == tuple(self.grid.shape[-2:])
        a = None if active is None else active.to(device=self.device, dtype=torch.uint8).contiguous()
        nat.check(self._lib.mapf_occupancy_accumulate(self._h, self._ptr(a), self._ptr(counts), self._stream()))
        return counts

    def distance_table(self) -> torch.Tensor:
        """uint8 [R*C, R*C]: moves between any two cells of the shared map around the obstacles (255 = unreachable),
        built once by a bit-parallel BFS kernel (one warp per source cell) and cached."""
        if getattr(self, "_dist_table", None) is None:
            cells = int(self.grid.shape[-2] * self.grid.shape[-1])
            t = torch.empty((cells, cells), dtype=torch.uint8, device=self.device)
            nat.check(self._lib.mapf_distance_table(self._h, self._ptr(t), self._stream()))
            self._dist_table = t
        return self._dist_table

    def goal_path_lengths(self) -> torch.Tensor:
        """int16 [B,N]: shortest obstacle-aware path length from every agent's position to its goal (-1 = unreachable),
        other agents ignored -- the single-agent lower bound the classical planners start from (scripts/a-star.py)."""
        out = torch.empty((self.B, self.N), dtype=torch.int16, device=self.device)
        nat.check(self._lib.mapf_goal_path_lengths(self._h, self._ptr(self.distance_table()), self._ptr(out), self._stream()))
        return out

    def poll_errors(self) -> int:
        bits = C.c_uint32(0)
        nat.check(self._lib.mapf_poll_errors(self._h, C.byref(bits), self._stream()))
        return int(bits.value)

    def raise_on_device_errors(self):
        """Raise the reference's exception types for conditions the kernels flagged."""
        bits = self.poll_errors()
        if bits & nat.DEV_ERR_INVALID_ACTION:
            raise ValueError("Invalid action (outside 0..4) in the batch")  # ENV:504-506
        if bits & nat.DEV_ERR_TOO_FEW_CELLS:
            raise ValueError("Environment has too few free cells for starts and goals")  # ENV:270-275
        if bits & nat.DEV_ERR_DUPLICATE_LAYOUT:
            raise ValueError("starts / goals override repeats a cell: a layout is 2N distinct cells (ENV:277); "
                             "inject co-located agents through set_state instead")
        if bits & nat.DEV_ERR_NO_GOAL_CELL:
            raise RuntimeError("No valid cell available for lifelong goal reassignment.")  # ENV:296-298

    @property
    def launch_count(self) -> int:
        return int(self._lib.mapf_launch_count(self._h))

    def get_state(self) -> dict:
        """Snapshot (clone) of the full env state; feeds :meth:`set_state` (checkpoint/resume)."""
        return {k: v.clone() for k, v in self.state.items()}

    def set_state(self, state: dict):
        for k, v in state.items():
            self.state[k].copy_(torch.as_tensor(v).to(self.state[k].dtype).reshape(self.state[k].shape))
