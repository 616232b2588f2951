#!/usr/bin/env python
"""Host-buffer step (mapf_step_host) under different transfer settings, all ranks of a node at once:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/e2e_probe.py

For each setting (plain copies / bit-packed with different thread counts, pinned or not) every rank builds a fresh
C3 batch and times the same calls; rank 0 prints one JSON line per setting (aggregate agent-steps/s, slowest rank)
plus the measured ceilings (pinned D2H DMA, host fill / copy rate) with every rank probing at the same moment."""
import ctypes as C
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
import bench  # noqa: E402
from dl_reference_models_b200 import _native as nat  # noqa: E402
from dl_reference_models_b200.batched_env import BatchedMapfEnv  # noqa: E402

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)


def rmax(x):
    if world == 1:
        return float(x)
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def rsum(x):
    if world == 1:
        return float(x)
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


class A:
    shape, envs, agents, sensor_range, replicas = "c3", None, None, 2, 1


args = A()
bench.apply_shape(args)
cfg, grid = bench.workload(args)
cfg["grid"] = grid
B, N, V = args.envs, args.agents, 5
ncores = len(os.sched_getaffinity(0))
if rank == 0:
    print(json.dumps({"world": world, "host_cores": ncores, "affinity": sorted(os.sched_getaffinity(0))[:64]}), flush=True)

host_out = {
    "local_obs": torch.empty((B, N, V, V), dtype=torch.uint8).pin_memory(),
    "action_mask": torch.empty((B, N, 5), dtype=torch.int8).pin_memory(),
    "goal_delta": torch.empty((B, N, 2), dtype=torch.float32).pin_memory(),
    "blocking_prev": torch.empty((B, N), dtype=torch.uint8).pin_memory(),
    "reward": torch.empty((B, N), dtype=torch.float32).pin_memory(),
    "terminated": torch.empty((B,), dtype=torch.uint8).pin_memory(),
    "truncated": torch.empty((B,), dtype=torch.uint8).pin_memory(),
}
cout = nat.MapfOutputs(**{k: v.data_ptr() for k, v in host_out.items()})
delivered = sum(v.numel() * v.element_size() for v in host_out.values())
gen = torch.Generator().manual_seed(999 + rank)
acts = [torch.randint(0, 5, (B, N), dtype=torch.int8, generator=gen).pin_memory() for _ in range(4)]
lib = nat.lib()

SETTINGS = [
    ("auto", {}),
    ("plain", {"MAPF_HOST_PACK": "0"}),
    ("packed", {"MAPF_HOST_PACK": "1"}),
    ("packed_nt", {"MAPF_HOST_PACK": "1", "MAPF_HOST_NT": "1"}),
    ("packed_nopin", {"MAPF_HOST_PACK": "1", "MAPF_HOST_PIN": "0"}),
    ("packed_t2", {"MAPF_HOST_PACK": "1", "MAPF_HOST_THREADS": "2"}),
    ("packed_t3", {"MAPF_HOST_PACK": "1", "MAPF_HOST_THREADS": "3"}),
    ("packed_t6", {"MAPF_HOST_PACK": "1", "MAPF_HOST_THREADS": "6"}),
    ("packed_t8", {"MAPF_HOST_PACK": "1", "MAPF_HOST_THREADS": "8"}),
]
if os.environ.get("E2E_SETTINGS"):
    want = os.environ["E2E_SETTINGS"].split(",")
    SETTINGS = [x for x in SETTINGS if x[0] in want]
KNOBS = ("MAPF_HOST_PACK", "MAPF_HOST_PIN", "MAPF_HOST_THREADS", "MAPF_HOST_SLICES", "MAPF_HOST_PLAN", "MAPF_HOST_NT")
for name, envv in SETTINGS:
    for k in KNOBS:
        os.environ.pop(k, None)
    os.environ.update(envv)
    env = BatchedMapfEnv(cfg, B, dev, env_id_base=rank * B)
    env.reset()
    torch.cuda.synchronize(dev)
    for i in range(18):
        barrier()
        nat.check(lib.mapf_step_host(env._h, C.c_void_p(acts[i % 4].data_ptr()), None, None, C.byref(cout), 1))
    reps = []
    for rep in range(3):
        barrier()
        t0 = time.perf_counter()
        for i in range(20):
            nat.check(lib.mapf_step_host(env._h, C.c_void_p(acts[i % 4].data_ptr()), None, None, C.byref(cout), 1))
        torch.cuda.synchronize(dev)
        reps.append(rmax(time.perf_counter() - t0))
    h2d, d2h = C.c_int64(0), C.c_int64(0)
    nat.check(lib.mapf_host_transfer_bytes(env._h, C.byref(h2d), C.byref(d2h)))
    th, fill, cp = C.c_int32(0), C.c_double(0.0), C.c_double(0.0)
    barrier()
    nat.check(lib.mapf_host_memory_probe(env._h, C.c_int64(delivered), C.byref(th), C.byref(fill), C.byref(cp)))
    fill_all, cp_all = rsum(fill.value), rsum(cp.value)
    dt = float(np.median(reps))
    if rank == 0:
        print(json.dumps({"setting": name, "agg_agent_steps_per_s": world * B * N * 20 / dt, "ms_per_step": dt / 20 * 1e3,
                          "d2h_bytes": int(d2h.value), "threads": th.value, "fill_gbs_aggregate": fill_all,
                          "copy_gbs_aggregate": cp_all, "reps_ms": [round(r / 20 * 1e3, 3) for r in reps]}), flush=True)
    if name == "auto":   # compact delivery next to it: records only, nothing expanded on the host
        rs_ = int(lib.mapf_packed_record_bytes(V * V))
        rec = torch.empty((B * N * rs_,), dtype=torch.uint8).pin_memory()
        small = {k: torch.empty_like(env.out[k], device="cpu").pin_memory() for k in ("blocking_prev", "terminated", "truncated")}
        cs = nat.MapfOutputs(**{k: (small[k].data_ptr() if k in small else None) for k in nat.OUTPUT_FIELDS})
        for i in range(3):
            nat.check(lib.mapf_step_host_records(env._h, C.c_void_p(acts[i % 4].data_ptr()), C.c_void_p(rec.data_ptr()), C.byref(cs), 1))
        rr = []
        for rep in range(3):
            barrier()
            t0 = time.perf_counter()
            for i in range(20):
                nat.check(lib.mapf_step_host_records(env._h, C.c_void_p(acts[i % 4].data_ptr()), C.c_void_p(rec.data_ptr()), C.byref(cs), 1))
            torch.cuda.synchronize(dev)
            rr.append(rmax(time.perf_counter() - t0))
        dtr = float(np.median(rr))
        if rank == 0:
            print(json.dumps({"setting": "records", "agg_agent_steps_per_s": world * B * N * 20 / dtr, "ms_per_step": dtr / 20 * 1e3,
                              "d2h_bytes": 0, "threads": 0, "fill_gbs_aggregate": 0.0, "copy_gbs_aggregate": 0.0}), flush=True)
    env.close()
    del env

# DMA ceilings, all ranks at once
n = 64 << 20
d = torch.empty(n, dtype=torch.uint8, device=dev)
h = torch.empty(n, dtype=torch.uint8).pin_memory()
res = {}
for nm, (src, dst) in (("d2h", (d, h)), ("h2d", (h, d))):
    dst.copy_(src, non_blocking=True)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(8):
        dst.copy_(src, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    res[nm + "_gbs_aggregate"] = rsum(8 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9)
if rank == 0:
    print(json.dumps(res), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
