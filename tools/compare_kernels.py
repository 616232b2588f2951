"""Two step kernels side by side on the same configuration (C3-like, 20-step episodes, Philox draws, in-launch resets,
fused masked sampler): prints, per step, which state / output tensors differ (count and first indices) instead of
stopping at the first difference -- the first-contact check of a new or changed kernel.
usage: [DBG_KIND=pair|env|lane] python tools/compare_kernels.py [B] [steps] [agents] [ref_kind]"""
import sys

import torch

sys.path.insert(0, ".")
from dl_reference_models_b200 import _native as nat  # noqa: E402
from dl_reference_models_b200 import maps  # noqa: E402
from dl_reference_models_b200.batched_env import BatchedMapfEnv  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096 + 7
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 60
NA = int(sys.argv[3]) if len(sys.argv) > 3 else 16
REF = sys.argv[4] if len(sys.argv) > 4 else "env"
OUT_KEYS = ("local_obs", "action_mask", "goal_delta", "blocking_prev", "reward", "terminated", "truncated",
            "step_flags", "agent_step_flags", "info")
cfg = {"num_agents": NA, "sensor_range": 2, "steps_per_episode": 20, "lifelong_mapf": True, "seed": 4242,
       "grid": maps.random_obstacle_grid(32, 32, 0.30, 2026, min_free=32)}
a = BatchedMapfEnv(dict(cfg, step_kernel=REF), B, "cuda:0")
b = BatchedMapfEnv(dict(cfg, step_kernel=__import__("os").environ.get("DBG_KIND", "pair")), B, "cuda:0")
print("kinds", nat.lib().mapf_step_kernel_kind(a._h), nat.lib().mapf_step_kernel_kind(b._h))
oa, ob = a.reset(), b.reset()
acts_a = a.sample_actions(masked=True)
acts_b = b.sample_actions(masked=True)
a.fuse_sampler("masked")
b.fuse_sampler("masked")
bad_steps = 0
for s in range(STEPS):
    oa = a.step(acts_a, auto_reset=True)
    ob = b.step(acts_b, auto_reset=True)
    torch.cuda.synchronize()
    diffs = []
    for k in a.state:
        x, y = a.state[k], b.state[k]
        if not torch.equal(x, y):
            nz = torch.nonzero(x != y)
            diffs.append(f"state.{k}: {nz.shape[0]} first {nz[:3].tolist()}")
    for k in OUT_KEYS:
        x, y = getattr(oa, k), getattr(ob, k)
        if not torch.equal(x, y):
            nz = torch.nonzero(x != y)
            i = tuple(nz[0].tolist())
            diffs.append(f"out.{k}: {nz.shape[0]} first {nz[:3].tolist()} ref={x[i].item()} got={y[i].item()}")
    if not torch.equal(acts_a, acts_b):
        diffs.append(f"next actions: {(acts_a != acts_b).sum().item()}")
    if diffs:
        bad_steps += 1
        print(f"step {s}:")
        for d in diffs:
            print("   ", d)
        if bad_steps >= 3:
            break
print("errors", a.poll_errors(), b.poll_errors())
print("RESULT", "OK" if bad_steps == 0 else "MISMATCH")
