#!/usr/bin/env python
"""Throughput benchmark of the MAPF reset()/step() hot path (BASELINE.json metric: agent-steps/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Workload (BASELINE config 3, SURVEY 8d "C3"): 65 536 envs x 16 agents PER GPU on a 32x32 map with
30 % i.i.d. obstacles (map seed 2026), sensor_range 2, lifelong goal resampling, lock metrics on,
256 steps per episode with the reset inside the step launch (like run_benchmark's `if done:
reset()`, scripts/benchmark_multi_agent_env.py:89-95), actions uniform over the valid mask.
A "step" is one pass of the hot path over the whole batch: ONE launch of the step kernel, which
also draws the next step's masked-uniform actions (the benchmark sampler fused in).  Envs shard independently across GPUs (weak scaling, no data-path collective); the one
NCCL all-reduce (episode/lock metric sums) runs once after the timed region.

Prints ONE JSON line (rank 0).  `value` = agent-steps/s with inputs resident in HBM; `e2e` = the
same metric through the C ABI's host-buffer entry point (mapf_step_host: pinned host actions in,
all observation/reward/done channels out, copies inside the timed region); `roofline` = the step
kernel against the measured HBM copy bandwidth; `cpu_baseline` = the CPU oracle (a C port of the
reference's Python env) on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))

METRIC = "agent_steps_per_sec"
UNIT = "agent-steps/s"
# mapf_step_kernel_kind() -> kernel symbol (include/mapf_b200.h)
def kernel_name(kind: int, num_agents: int, sensor_range: int) -> str:
    if kind == 2:
        return f"mapf_step_env_kernel<{sensor_range},...> (env-per-thread)"
    g = 4 if num_agents <= 4 else 8 if num_agents <= 8 else 16 if num_agents <= 16 else 32
    return f"mapf_step_kernel<{g},{sensor_range},...> (lane-per-agent)"


def default_traffic(kind: int, shape: str = "c3"):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the step kernel at the named shape's default
    size, from the committed `ncu --set full` captures (profiles/README.md); None where there is no capture."""
    return {("c3", 1): 54.3e6, ("c3", 2): 75.3e6, ("c4", 1): 140.8e6}.get((shape, kind))


SHAPES = {
    # BASELINE.json configs[2] (the one `metric` is quoted on): the default
    "c3": {"envs": 65536, "agents": 16, "lifelong": True, "steps_per_episode": 256, "map": "32x32",
           "what": "32x32 map (30% obstacles, seed 2026), lifelong goal resampling"},
    # configs[1]: the parity-replay shape, reported for throughput too
    "c2": {"envs": 4096, "agents": 4, "lifelong": False, "steps_per_episode": 100, "map": "10x20",
           "what": "ReferenceModel-2-1 map (10x20), episodic (terminate when all agents are home)"},
    # configs[3]: deadlock-heavy corridors
    "c4": {"envs": 65536, "agents": 32, "lifelong": False, "steps_per_episode": 256, "map": "32x32 corridors",
           "what": "32x32 map of 1-wide corridors, episodic, deadlock-heavy"},
}


def apply_shape(args):
    sh = SHAPES[args.shape]
    if args.envs is None:
        args.envs = sh["envs"]
    if args.agents is None:
        args.agents = sh["agents"]


def workload(args) -> tuple[dict, np.ndarray]:
    from dl_reference_models_b200 import maps

    sh = SHAPES[args.shape]
    if args.shape == "c2":
        grid = maps.get_grid("ReferenceModel-2-1")
    elif args.shape == "c4":
        grid = maps.corridor_grid(32, 32)
    else:
        grid = maps.random_obstacle_grid(32, 32, 0.30, 2026, min_free=2 * args.agents)
    cfg = {
        "num_agents": args.agents, "sensor_range": args.sensor_range, "steps_per_episode": sh["steps_per_episode"],
        "lifelong_mapf": sh["lifelong"], "enable_lock_metrics": True, "deterministic": False, "seed": 999,
        "deadlock_window_steps": 8, "livelock_window_steps": 16,
    }
    return cfg, grid


def config_dict(args, n_gpus: int) -> dict:
    return {
        "workload": f"{args.shape.upper()}: {args.envs} envs x {args.agents} agents per GPU, {SHAPES[args.shape]['what']}, "
                    f"sensor_range {args.sensor_range}, lock metrics on, "
                    f"{SHAPES[args.shape]['steps_per_episode']} steps/episode, in-launch auto-reset, "
                    "masked-uniform actions sampled on device",
        "envs_per_gpu": args.envs, "num_agents": args.agents, "map": SHAPES[args.shape]["map"],
        "sensor_range": args.sensor_range,
        "sharding": f"envs x{n_gpus} (independent shards, no data-path collective)",
        "l2": f"{args.replicas} rotating replicas of the batch (working set > 126 MB L2 between launches)",
    }


def algorithmic_bytes_per_agent_step(N: int, V: int, lock: bool = True, map_bytes: int = 0) -> float:
    """33 + V^2 + 28*[lock] + (60 + map_bytes)/N."""
    return 33 + V * V + (28 if lock else 0) + (60 + map_bytes) / N


def measured_peaks() -> tuple[float, str]:
    p = REPO / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.t0, self.t1 = index, [], None, None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-i", str(self.index),
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def wait_first_sample(self, timeout: float = 10.0):
        t = time.perf_counter()
        while self.proc is not None and not self.rows and time.perf_counter() - t < timeout:
            time.sleep(0.01)

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.thread.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ts, r in self.rows:
            # a sample describes the interval before it: keep those taken under load
            if len(r) < 6 or self.t0 is None or ts < self.t0 + 0.02 or ts > self.t1 + 0.03:
                continue
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except ValueError:
                continue
            for name, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


class CpuPort:
    """The CPU oracle (C port of the reference's Python env) on all host cores, timed on bounded
    samples of the bench workload.  bench.py is one of the few places allowed to execute oracle/."""

    def __init__(self, args, threads: int | None = None):
        from oracle import oracle as orc

        self.orc, self.args = orc, args
        self.cfg, self.grid = workload(args)
        self.threads = threads or len(os.sched_getaffinity(0))
        self.envs = orc.bench_envs(self.cfg, self.grid, self.threads * 32)
        n, _, dt, _ = self._run(20)  # calibration: env-steps/s
        self.rate = n / max(dt, 1e-9)

    def _run(self, steps: int):
        return self.orc.bench_run(self.cfg, self.grid, len(self.envs), steps, mode="masked", threads=self.threads,
                                  envs=self.envs)

    def sample(self, seconds: float) -> dict:
        steps = int(max(4, min(1_000_000, seconds * self.rate / len(self.envs))))
        n, episodes, dt, used = self._run(steps)
        self.rate = n / max(dt, 1e-9)
        return {
            "value": n * self.args.agents / dt, "unit": UNIT, "cores": used, "kind": "port",
            "sample": f"{len(self.envs)} envs x {steps} steps ({n} env-steps, {episodes} episodes, {dt:.2f} s) of the "
                      f"same workload on {used} host threads; C port of the reference's Python env "
                      "(oracle/mapf_oracle.c)",
        }


def cpu_port_run(args, seconds: float = 12.0) -> dict:
    return CpuPort(args).sample(seconds)


def run_reference(args, rank: int, world: int):
    """--impl reference: the reference's CPU implementation of the path on the host cores.  The
    reference is pure Python and cannot travel to the GPU box (it is not vendored), so this arm
    times its C port, the oracle -- a generous stand-in (the Python original is ~100x slower)."""
    if rank != 0:
        return
    total = args.warmup + args.steps
    per = min(15.0, 100.0 / max(1, total))
    port = CpuPort(args)
    vals, last = [], None
    t0 = time.perf_counter()
    for i in range(total):
        last = port.sample(per)
        if i >= args.warmup:
            vals.append(last["value"])
    wall = time.perf_counter() - t0
    v = float(np.mean(vals))
    cb = {"value": v, "unit": UNIT, "cores": last["cores"], "kind": "port",
          "sample": f"mean of {len(vals)} samples ({wall:.0f} s in total), the last one: " + last["sample"]}
    line = {
        "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": None, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16/u8",
        "data": "synthetic", "impl": "reference", "config": config_dict(args, args.gpus), "cpu_baseline": cb,
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_b200(args, rank: int, local_rank: int, world: int):
    import torch
    import torch.distributed as dist

    from dl_reference_models_b200 import _native as nat
    from dl_reference_models_b200.batched_env import BatchedMapfEnv
    from dl_reference_models_b200.metrics import allreduce_metrics

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (this framework has no CPU fallback; use --impl reference "
                         "for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg, grid = workload(args)
    cfg["grid"] = grid
    B, N, V = args.envs, args.agents, 2 * args.sensor_range + 1
    envs = [BatchedMapfEnv(cfg, B, dev, env_id_base=(rank * args.replicas + r) * B) for r in range(args.replicas)]
    for e in envs:
        e.reset()
    K, W = args.steps, args.warmup

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # the benchmark's masked action sampler is fused into the step launch: one kernel per env step
    for e in envs:
        e._next = e.sample_actions(masked=True)
        e.fuse_sampler("masked")

    for i in range(W):
        e = envs[i % len(envs)]
        e.step(e._next, auto_reset=True)
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = sum(e.launch_count for e in envs)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        sampler.wait_first_sample()
    barrier()
    sampler.mark_begin()
    t_begin.record()
    for i in range(K):
        e = envs[(W + i) % len(envs)]
        e.step(e._next, auto_reset=True)
    t_end.record()
    barrier()
    sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None
    launches = sum(e.launch_count for e in envs) - launches0
    kind = int(nat.lib().mapf_step_kernel_kind(envs[0]._h))
    ms_total = t_begin.elapsed_time(t_end)
    # one step = ONE launch of the step kernel (checked below), so the kernel's average launch duration is the
    # CUDA-event time of the region / K -- launch gaps included; no events between the launches (they would
    # serialise the stream and keep the next launch's ramp-up from overlapping this launch's tail)
    kernel_ms = ms_total / K
    assert launches == K, (launches, K)
    for e in envs:
        e.raise_on_device_errors()

    # ---- e2e: the C ABI's host-buffer entry point, pinned host memory, copies inside the timed region
    import ctypes as C
    e2e_env = envs[0]
    e2e_env.fuse_sampler(None)
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
    gen = torch.Generator().manual_seed(999 + rank)
    host_actions = [torch.randint(0, 5, (B, N), dtype=torch.int8, generator=gen).pin_memory() for _ in range(4)]
    delivered = sum(v.numel() * v.element_size() for v in host_out.values())
    lib = nat.lib()
    Ke, We = max(3, min(K, 50)), 3
    torch.cuda.synchronize(dev)
    for i in range(We):
        nat.check(lib.mapf_step_host(e2e_env._h, C.c_void_p(host_actions[i % 4].data_ptr()), None, None,
                                     C.byref(cout), 1))
    barrier()
    t0 = time.perf_counter()
    for i in range(Ke):
        nat.check(lib.mapf_step_host(e2e_env._h, C.c_void_p(host_actions[i % 4].data_ptr()), None, None,
                                     C.byref(cout), 1))
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    checksum = float(host_out["reward"].sum())  # the step's result is really on the host
    # bytes that crossed PCIe in one call (the big per-agent channels travel bit-packed and are expanded into the
    # host arrays by the library's host threads inside the call; `delivered` is the size of the arrays filled)
    c_h2d, c_d2h = C.c_int64(0), C.c_int64(0)
    nat.check(lib.mapf_host_transfer_bytes(e2e_env._h, C.byref(c_h2d), C.byref(c_d2h)))
    h2d, d2h = int(c_h2d.value), int(c_d2h.value)

    # ---- reduce over ranks (max time), metric all-reduce off the step path
    t = torch.tensor([ms_total, kernel_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    nl = torch.tensor([float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(nl, op=dist.ReduceOp.SUM)
    ms_total, kernel_ms, e2e_ms = (float(x) for x in t.tolist())
    launches = int(nl.item())
    mvec = envs[0].metrics_vector().clone()
    for e in envs[1:]:
        mvec += e.metrics_vector()
    metrics = allreduce_metrics(mvec, world)

    if rank == 0:
        peak, peak_src = measured_peaks()
        bytes_per_as = algorithmic_bytes_per_agent_step(N, V, True, 0)
        alg_bytes = bytes_per_as * B * N
        achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
        value = world * B * N * K / (ms_total * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int16/u8", "data": "synthetic", "impl": "b200", "config": config_dict(args, world),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": args.traffic_bytes if args.traffic_bytes is not None else (
                             default_traffic(kind, args.shape) if (B, N, V) == (65536, SHAPES[args.shape]["agents"], 5) else None),
                         "kernel": kernel_name(kind, N, args.sensor_range), "kernel_ms": kernel_ms,
                         "algorithmic_bytes_per_launch": alg_bytes, "bytes_per_agent_step": bytes_per_as,
                         "peak_source": peak_src},
            "e2e": {"value": world * B * N * Ke / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "delivered_bytes_per_step": delivered, "steps": Ke,
                    "api": "mapf_step_host (C ABI, pinned host buffers)",
                    "transfer": ("bit-packed agent records over PCIe, expanded into the host arrays inside the call"
                                 if d2h < delivered else "plain copies"),
                    "actions": "uniform random from pinned host buffers", "checksum": checksum},
            "gpu_launches": int(launches), "clocks": clocks, "step_kernel": {1: "lane", 2: "env"}.get(kind),
            "episode_metrics": {k: metrics[k] for k in ("episodes", "goals_reached_mean", "deadlock_steps_mean",
                                                        "livelock_steps_mean", "throughput_mean")},
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_port_run(args, seconds=args.cpu_seconds)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4000)
    ap.add_argument("--warmup", type=int, default=100)
    ap.add_argument("--impl", choices=("b200", "reference"), default="b200")
    ap.add_argument("--shape", choices=tuple(SHAPES), default="c3",
                    help="named shape of BASELINE.json (default c3: the one the metric is quoted on)")
    ap.add_argument("--envs", type=int, default=None, help="envs per GPU (default: the shape's)")
    ap.add_argument("--agents", type=int, default=None)
    ap.add_argument("--sensor-range", type=int, default=2)
    ap.add_argument("--replicas", type=int, default=4, help="independent batches rotated between launches (L2 defeat)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--traffic-bytes", type=float, default=None,
                    help="dram bytes per step-kernel launch from the committed ncu --set full capture")
    args = ap.parse_args()
    apply_shape(args)
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world == 1 and args.gpus > 1:
        # convenience: re-launch under torchrun, one rank per GPU
        import socket
        s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(port), str(Path(__file__).resolve())] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    run_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
