#!/usr/bin/env python
"""Kernel A/B timing of one libmapf_b200.so build (variant libraries from tools/build_variant.sh):

    MAPF_B200_LIB=scratch/variants/libmapf_X.so python tools/bench_kernel.py --shape c3 --steps 400

Steady state like bench.py: episode phases staggered over the episode length, burn-in, 4 rotating replicas (L2 defeat),
fused masked sampler, in-launch auto-reset; CUDA events around the timed launches.  Prints one JSON line."""
import argparse
import ctypes
import json
import os
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="c3")
ap.add_argument("--steps", type=int, default=400)
ap.add_argument("--burn", type=int, default=48)
ap.add_argument("--envs", type=int, default=None)
ap.add_argument("--agents", type=int, default=None)
ap.add_argument("--sensor-range", type=int, default=2)
ap.add_argument("--replicas", type=int, default=4)
ap.add_argument("--no-stagger", action="store_true")
ap.add_argument("--blocks", type=int, default=5)
ap.add_argument("--inner", type=int, default=1, help="env steps per launch (mapf_step_many, rollout buffers [K, ...])")
ap.add_argument("--episode-steps", type=int, default=None, help="override steps_per_episode (reset frequency experiments)")
args = ap.parse_args()

import torch  # noqa: E402

import bench  # noqa: E402
from dl_reference_models_b200 import _native as nat  # noqa: E402

# older variant libraries may lack the newest C-ABI symbols: alias them so that the binding loads (never called here)
_L = ctypes.CDLL(str(nat.LIB_PATH))
_missing = [s for s in nat.EXPORTS if not hasattr(_L, s)]
if _missing:
    _real = ctypes.CDLL

    class _Shim:
        def __init__(self, lib):
            object.__setattr__(self, "_lib", lib)

        def __getattr__(self, name):
            try:
                return getattr(self._lib, name)
            except AttributeError:
                return self._lib.mapf_launch_count

    nat.C.CDLL = lambda p: _Shim(_real(p))
from dl_reference_models_b200.batched_env import BatchedMapfEnv  # noqa: E402

bench.apply_shape(args)
cfg, grid = bench.workload(args)
cfg["grid"] = grid
if args.episode_steps:
    cfg["steps_per_episode"] = args.episode_steps
dev = torch.device("cuda", 0)
envs = [BatchedMapfEnv(cfg, args.envs, dev, env_id_base=r * args.envs) for r in range(args.replicas)]
T = int(cfg["steps_per_episode"])
for e in envs:
    e.reset()
    if not args.no_stagger:
        g = torch.Generator().manual_seed(2026 + envs.index(e))
        e.state["env_words"][:, nat.W_STEP_COUNT] = (torch.randperm(args.envs, generator=g) % T).to(device=dev, dtype=torch.int32)
    e._next = e.sample_actions(masked=True)
    e.fuse_sampler("masked")
for i in range(args.burn * len(envs)):
    e = envs[i % len(envs)]
    e.step(e._next, auto_reset=True)
if args.inner > 1:
    for e in envs:
        e._roll = e.rollout_buffers(args.inner)
torch.cuda.synchronize()
times = []
for b in range(args.blocks):
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    if args.inner > 1:
        for i in range(max(1, args.steps // args.inner)):
            e = envs[i % len(envs)]
            e.step_many(args.inner, out=e._roll)
    else:
        for i in range(args.steps):
            e = envs[i % len(envs)]
            e.step(e._next, auto_reset=True)
    t1.record()
    torch.cuda.synchronize()
    nsteps = max(1, args.steps // args.inner) * args.inner if args.inner > 1 else args.steps
    times.append(t0.elapsed_time(t1) * 1e3 / nsteps)
for e in envs:
    e.raise_on_device_errors()
eps = sum(float(e.metrics_vector()[0]) for e in envs)
kind = int(nat.lib().mapf_step_kernel_kind(envs[0]._h))
print(json.dumps({"lib": os.path.basename(str(nat.LIB_PATH)), "shape": args.shape, "envs": args.envs, "agents": args.agents,
                  "kernel": {1: "lane", 2: "env", 3: "pair"}[kind], "episode_steps": T, "inner": args.inner, "us_per_step_median": sorted(times)[len(times) // 2],
                  "us_per_step_blocks": [round(t, 2) for t in times], "episodes": eps}))
