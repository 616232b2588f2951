"""The two step kernels -- lane-per-agent (one agent per lane, north_star's mapping) and env-per-thread
(one env per thread, bitboards in shared memory) -- are the same function: identical state tensors and
identical outputs, step by step, including Philox goal draws, in-launch auto-resets and the fused sampler.
The oracle / golden parity suites run on whichever kernel `auto` selects; this file ties the other one to it."""
import numpy as np
import pytest

from dl_reference_models_b200 import _native as nat

pytestmark = pytest.mark.gpu

OUT_KEYS = ("local_obs", "action_mask", "goal_delta", "blocking_prev", "reward", "terminated", "truncated",
            "step_flags", "agent_step_flags", "info")


def make(cfg, B, kind, **kw):
    from dl_reference_models_b200.batched_env import BatchedMapfEnv

    env = BatchedMapfEnv(dict(cfg, step_kernel=kind), B, "cuda:0", **kw)
    assert nat.lib().mapf_step_kernel_kind(env._h) == {"lane": 1, "env": 2, "pair": 3}[kind]
    return env


def assert_same(a, b, oa, ob, ctx):
    import torch

    for k in a.state:
        if k == "lock_distance":
            continue  # compared below through the window the kernels read (unused slots may differ after a reset)
        assert torch.equal(a.state[k], b.state[k]), f"{ctx}: state {k} differs at {torch.nonzero(a.state[k] != b.state[k])[:4].tolist()}"
    assert torch.equal(a.state["lock_distance"], b.state["lock_distance"]), f"{ctx}: lock_distance"
    for k in OUT_KEYS:
        x, y = getattr(oa, k), getattr(ob, k)
        assert torch.equal(x, y), f"{ctx}: output {k} differs at {torch.nonzero(x != y)[:4].tolist()}"


def run_pair(cfg, B, steps, masked=True, auto_reset=True, fused=True, kinds=("lane", "env")):
    import torch

    a, b = make(cfg, B, kinds[0]), make(cfg, B, kinds[1])
    oa, ob = a.reset(), b.reset()
    assert_same(a, b, oa, ob, "reset")
    acts_a = a.sample_actions(masked=masked)
    acts_b = b.sample_actions(masked=masked)
    if fused:
        a.fuse_sampler("masked" if masked else "random")
        b.fuse_sampler("masked" if masked else "random")
    for s in range(steps):
        if not fused:
            acts_a = a.sample_actions(masked=masked)
            acts_b = b.sample_actions(masked=masked)
        assert torch.equal(acts_a, acts_b), f"step {s}: sampled actions differ"
        oa = a.step(acts_a, auto_reset=auto_reset)
        ob = b.step(acts_b, auto_reset=auto_reset)
        assert_same(a, b, oa, ob, f"step {s}")
    assert a.poll_errors() == b.poll_errors()
    assert np.array_equal(a.metrics_vector().cpu().numpy(), b.metrics_vector().cpu().numpy())
    return a, b


def c3(**kw):
    from dl_reference_models_b200 import maps

    cfg = {"num_agents": 16, "sensor_range": 2, "steps_per_episode": 40, "lifelong_mapf": True, "seed": 4242,
           "grid": maps.random_obstacle_grid(32, 32, 0.30, 2026, min_free=32)}
    cfg.update(kw)
    return cfg


def test_c3_lifelong_autoreset_fused_sampler():
    a, _ = run_pair(c3(), 4096 + 7, 100)
    assert int(a.state["env_words"][:, nat.W_EPISODES].min()) == 2


@pytest.mark.parametrize("sr", [1, 3])
def test_sensor_ranges(sr):
    run_pair(c3(sensor_range=sr, steps_per_episode=25), 700, 60)


@pytest.mark.parametrize("n", [2, 5, 7, 12, 32])
def test_agent_counts_including_non_multiples_of_four(n):
    run_pair(c3(num_agents=n, steps_per_episode=30), 333, 70)


def test_crowded_random_actions_many_failed_moves_and_goal_draws():
    """Dense crowd + unmasked actions: many blocked moves, follow chains, wait-for cycles, goal draws."""
    from dl_reference_models_b200 import maps

    cfg = c3(num_agents=32, grid=maps.random_obstacle_grid(12, 12, 0.15, 7, min_free=80), steps_per_episode=50)
    a, _ = run_pair(cfg, 1000, 120, masked=False)
    assert int(a.state["env_words"][:, nat.W_WFG_CYCLE_STEPS].max()) >= 0


def test_non_lifelong_reference_map_deterministic_and_random():
    for det in (True, False):
        cfg = {"env_name": "ReferenceModel-2-1", "num_agents": 4, "sensor_range": 2, "steps_per_episode": 100,
               "deterministic": det, "seed": 123}
        run_pair(cfg, 513, 230, masked=True)
        run_pair(cfg, 129, 120, masked=False, fused=False)


def test_corridor_deadlocks_lock_metrics():
    """BASELINE config 4 shape: 1-wide corridors, 32 agents, non-lifelong, dw=8 / lw=16."""
    from dl_reference_models_b200 import maps

    cfg = {"num_agents": 32, "sensor_range": 2, "steps_per_episode": 256, "lifelong_mapf": False, "seed": 77,
           "deadlock_window_steps": 8, "livelock_window_steps": 16, "grid": maps.corridor_grid(32, 32)}
    a, _ = run_pair(cfg, 256, 300, masked=False)
    w = a.state["env_words"].cpu().numpy()
    m = a.metrics_vector().cpu().numpy()
    assert m[nat.METRIC_NAMES.index("deadlock_steps_sum")] + w[:, nat.W_DEADLOCK_STEPS].sum() > 0


def test_small_maps_and_lock_window_variants():
    for name, n in (("ReferenceModel-1-4", 4), ("ReferenceModel-1-2", 2), ("ReferenceModel-3-1", 8)):
        cfg = {"env_name": name, "num_agents": n, "sensor_range": 2, "steps_per_episode": 60, "lifelong_mapf": True,
               "deadlock_window_steps": 2, "livelock_window_steps": 4, "lock_nearby_manhattan": 3, "seed": 5}
        run_pair(cfg, 200, 150, masked=False)


def test_sliced_host_step_equals_device_step(monkeypatch):
    """mapf_step_host pipelines big batches in slices over two streams (H2D / kernel / D2H overlap); slicing must not
    change anything: same outputs as one device-side step of an identical env, for 1, 3 and the default 2 slices."""
    import ctypes as C

    import torch

    B = 8192 + 64
    cfg = c3(steps_per_episode=12)
    for slices in ("1", "3", None):
        if slices is None:
            monkeypatch.delenv("MAPF_HOST_SLICES", raising=False)
        else:
            monkeypatch.setenv("MAPF_HOST_SLICES", slices)
        a, b = make(cfg, B, "lane"), make(cfg, B, "lane")
        a.reset()
        b.reset()
        host = {k: torch.empty_like(v, device="cpu").pin_memory() for k, v in b.out.items()}
        cout = nat.MapfOutputs(**{k: host[k].data_ptr() for k in nat.OUTPUT_FIELDS})
        gen = torch.Generator().manual_seed(5)
        for s in range(30):
            acts = torch.randint(0, 5, (B, a.N), dtype=torch.int8, generator=gen)
            oa = a.step(acts.cuda(), auto_reset=True)
            nat.check(nat.lib().mapf_step_host(b._h, C.c_void_p(acts.pin_memory().data_ptr()), None, None, C.byref(cout), 1))
            for k in OUT_KEYS:
                assert torch.equal(getattr(oa, k).cpu(), host[k]), f"slices={slices} step {s}: {k}"
        for k in a.state:
            assert torch.equal(a.state[k], b.state[k]), f"slices={slices}: state {k}"




def test_tall_map_null_actions_and_invalid_actions():
    """Edge inputs on both kernels: a 48-row map (boards taller than a warp), NULL actions (= all NO_OP, ENV:498-500),
    out-of-range actions (flagged MAPF_DEV_ERR_INVALID_ACTION and treated as NO_OP), maximum agent count."""
    import torch

    from dl_reference_models_b200 import maps

    grid = maps.random_obstacle_grid(48, 32, 0.25, 11, min_free=200)
    cfg = {"num_agents": 32, "sensor_range": 3, "steps_per_episode": 20, "lifelong_mapf": True, "seed": 3, "grid": grid}
    a, b = make(cfg, 257, "lane"), make(cfg, 257, "env")
    oa, ob = a.reset(), b.reset()
    assert_same(a, b, oa, ob, "reset")
    gen = torch.Generator().manual_seed(0)
    for s in range(50):
        if s % 7 == 3:
            oa, ob = a.step(None, auto_reset=True), b.step(None, auto_reset=True)
            assert int(oa.agent_step_flags.bitwise_and(nat.ASF_MOVED).sum()) == 0, "NULL actions move nobody"
        else:
            acts = torch.randint(-2 if s % 11 == 5 else 0, 7 if s % 11 == 5 else 5, (257, 32), dtype=torch.int8, generator=gen)
            oa, ob = a.step(acts, auto_reset=True), b.step(acts, auto_reset=True)
        assert_same(a, b, oa, ob, f"step {s}")
    ea, eb = a.poll_errors(), b.poll_errors()
    assert ea == eb and (ea & nat.DEV_ERR_INVALID_ACTION)


def test_generic_instantiation_on_the_fast_configuration(monkeypatch):
    """Lifelong + lock metrics normally run the FAST instantiation of the env-per-thread kernel (those two switches
    compile-time); MAPF_ENV_FAST=0 keeps the generic one on the same configuration -- still the same function."""
    monkeypatch.setenv("MAPF_ENV_FAST", "0")
    run_pair(c3(steps_per_episode=25), 2048 + 5, 60)


@pytest.mark.parametrize("n", [4, 12])
def test_full_quads_other_agent_counts(n):
    """N % 4 == 0 takes the quad path without per-agent bound tests (and, lifelong + lock metrics, the FAST
    instantiation): 1 and 3 quads per env, besides the 2 and 4 of the other tests."""
    run_pair(c3(num_agents=n, steps_per_episode=30), 1024 + 9, 70)
    run_pair(c3(num_agents=n, steps_per_episode=30, lifelong_mapf=False), 512 + 3, 70, masked=False)


def test_two_lanes_per_env_kernel_equals_the_other_two():
    """step_kernel="pair" (mapf_pair_kernel.cuh: two lanes per env, the moves of a quad serialised half by half, shared
    64-bit owner boards): the same function as the env-per-thread and the lane-per-agent kernels -- lifelong and
    episodic, 1 to 8 quads per env, every sensor range, both instantiations, corridors with lock events, injected
    invalid / NULL actions on a tall map."""
    import torch

    from dl_reference_models_b200 import maps

    a, _ = run_pair(c3(), 4096 + 7, 100, kinds=("env", "pair"))
    assert int(a.state["env_words"][:, nat.W_EPISODES].min()) == 2
    for n in (4, 8, 12):
        run_pair(c3(num_agents=n, steps_per_episode=30), 1024 + 9, 70, kinds=("env", "pair"))
        run_pair(c3(num_agents=n, steps_per_episode=30, lifelong_mapf=False), 512 + 3, 70, masked=False, kinds=("lane", "pair"))
    for sr in (1, 3):
        run_pair(c3(sensor_range=sr, steps_per_episode=25), 700, 60, kinds=("env", "pair"))
    cfg = {"num_agents": 32, "sensor_range": 2, "steps_per_episode": 256, "lifelong_mapf": False, "seed": 77,
           "deadlock_window_steps": 8, "livelock_window_steps": 16, "grid": maps.corridor_grid(32, 32)}
    run_pair(cfg, 256, 300, masked=False, kinds=("lane", "pair"))
    cfg = c3(num_agents=32, grid=maps.random_obstacle_grid(12, 12, 0.15, 7, min_free=80), steps_per_episode=50)
    run_pair(cfg, 1000, 120, masked=False, kinds=("env", "pair"))
    for name, n in (("ReferenceModel-1-4", 4), ("ReferenceModel-3-1", 8)):
        cfg = {"env_name": name, "num_agents": n, "sensor_range": 2, "steps_per_episode": 60, "lifelong_mapf": True,
               "deadlock_window_steps": 2, "livelock_window_steps": 4, "lock_nearby_manhattan": 3, "seed": 5}
        run_pair(cfg, 200, 150, masked=False, kinds=("lane", "pair"))
    grid = maps.random_obstacle_grid(48, 32, 0.25, 11, min_free=200)
    cfg = {"num_agents": 32, "sensor_range": 3, "steps_per_episode": 20, "lifelong_mapf": True, "seed": 3, "grid": grid}
    a, b = make(cfg, 257, "env"), make(cfg, 257, "pair")
    oa, ob = a.reset(), b.reset()
    gen = torch.Generator().manual_seed(0)
    for s in range(40):
        if s % 7 == 3:
            oa, ob = a.step(None, auto_reset=True), b.step(None, auto_reset=True)
        else:
            acts = torch.randint(-2 if s % 11 == 5 else 0, 7 if s % 11 == 5 else 5, (257, 32), dtype=torch.int8, generator=gen)
            oa, ob = a.step(acts, auto_reset=True), b.step(acts, auto_reset=True)
        assert_same(a, b, oa, ob, f"step {s}")
    assert a.poll_errors() == b.poll_errors()


@pytest.mark.parametrize("lifelong", [False, True])
def test_pair_kernel_injected_colocation_equals_env_kernel(lifelong):
    """States no legal step produces (several agents on one cell, ENV:658-666): the env-per-thread kernel replays the
    live reference's traces of such states (tests/test_gpu_colocation.py); the two-lanes-per-env kernel keeps the same
    owner-grid semantics -- same penalties, same centre cells, same NOT_OWNER flags, moving co-located agents included."""
    import torch

    from dl_reference_models_b200 import maps

    B, n = 300, 8
    cfg = {"num_agents": n, "sensor_range": 2, "steps_per_episode": 12, "lifelong_mapf": lifelong, "seed": 9,
           "deadlock_window_steps": 2, "livelock_window_steps": 4, "grid": maps.random_obstacle_grid(9, 11, 0.2, 3, min_free=40)}
    a, b = make(cfg, B, "env"), make(cfg, B, "pair")
    a.reset(); b.reset()
    gen = torch.Generator().manual_seed(1)
    st = a.get_state()
    pos = st["positions"].cpu().clone()
    src = torch.randint(0, n, (B, 3), generator=gen)
    dst = torch.randint(0, n, (B, 3), generator=gen)
    for j in range(3):   # agent dst[b, j] is put on the cell of agent src[b, j]: pairs, triples and chains across both halves
        idx = torch.arange(B)
        pos[idx, dst[:, j]] = pos[idx, src[:, j]].clone()
    st["positions"] = pos
    a.set_state(st)
    b.set_state({k: v.clone() for k, v in st.items()})
    for s in range(10):
        acts = torch.randint(0, 5, (B, n), dtype=torch.int8, generator=gen)
        oa, ob = a.step(acts, auto_reset=False), b.step(acts, auto_reset=False)
        assert_same(a, b, oa, ob, f"step {s}")
    assert float(oa.reward.min()) <= 0.0


def test_pair_kernel_generic_instantiation(monkeypatch):
    monkeypatch.setenv("MAPF_ENV_FAST", "0")
    run_pair(c3(steps_per_episode=25), 2048 + 5, 60, kinds=("env", "pair"))


def test_pair_kernel_needs_full_quads():
    with pytest.raises(Exception):
        make(c3(num_agents=5), 64, "pair")


@pytest.mark.parametrize("kind,B,n", [("lane", 4096, 4), ("lane", 300, 16), ("env", 2048, 8), ("env", 4096 + 33, 16),
                                      ("env", 515, 5), ("pair", 1024, 16)])
def test_step_many_equals_single_steps(kind, B, n):
    """mapf_step_many: K env steps with the fused sampler's actions in ONE launch (both the lane-per-agent and the
    env-per-thread kernel -- every warp takes its tile through the K steps; K launches behind the same call for the
    two-lanes-per-env kernel) -- every output row of every step and the state afterwards equal K single mapf_step
    calls, bit for bit, across in-launch resets."""
    import torch

    K = 24
    cfg = c3(num_agents=n, steps_per_episode=10) if n != 4 else {
        "env_name": "ReferenceModel-2-1", "num_agents": 4, "sensor_range": 2, "steps_per_episode": 9, "seed": 11}
    a, b = make(cfg, B, kind), make(cfg, B, kind)
    a.reset(); b.reset()
    for e in (a, b):
        e._next = e.sample_actions(masked=True)
        e.fuse_sampler("masked")
    rows = {k: [] for k in nat.OUTPUT_FIELDS}
    for _ in range(K):
        a.step(a._next, auto_reset=True)
        for k in nat.OUTPUT_FIELDS:
            rows[k].append(a.out[k].clone())
    buf = b.rollout_buffers(K)
    b.step_many(K, out=buf)
    for k in nat.OUTPUT_FIELDS:
        assert torch.equal(torch.stack(rows[k]), buf[k]), k
    for k in a.state:
        assert torch.equal(a.state[k], b.state[k]), k
    assert torch.equal(a._actions, b._actions)
    # and again without a rollout buffer: the env's own buffers hold the last step
    for _ in range(5):
        a.step(a._next, auto_reset=True)
    b.step_many(5)
    for k in nat.OUTPUT_FIELDS:
        assert torch.equal(a.out[k], b.out[k]), k
    assert a.poll_errors() == 0 and b.poll_errors() == 0
