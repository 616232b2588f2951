"""Size-independent properties of the CUDA path: BASELINE config 3 at full size (65 536 envs x 16
agents), the Philox reset / goal draws (distributional equivalence with the reference's uniform
draws), shard invariance, the fused sampler, device error flags and the wait-for-graph output."""
import numpy as np
import pytest

from dl_reference_models_b200 import _native as nat

pytestmark = pytest.mark.gpu


def make(cfg, B, **kw):
    from dl_reference_models_b200.batched_env import BatchedMapfEnv

    return BatchedMapfEnv(cfg, B, "cuda:0", **kw)


def c3_cfg(**kw):
    from dl_reference_models_b200 import maps

    cfg = {"num_agents": 16, "sensor_range": 2, "steps_per_episode": 64, "lifelong_mapf": True, "seed": 999,
           "grid": maps.random_obstacle_grid(32, 32, 0.30, 2026, min_free=32)}
    cfg.update(kw)
    return cfg


def check_invariants(env, out, lifelong):
    import torch

    grid = torch.from_numpy(env.grid.astype(np.int64)).to(env.device)
    R, C = env.grid.shape
    pos = env.state["positions"].long()
    goals = env.state["goals"].long()
    assert ((pos[..., 0] >= 0) & (pos[..., 0] < R) & (pos[..., 1] >= 0) & (pos[..., 1] < C)).all()
    assert (grid[pos[..., 0], pos[..., 1]] == 0).all(), "agent on an obstacle"
    assert (grid[goals[..., 0], goals[..., 1]] == 0).all(), "goal on an obstacle"
    lin = pos[..., 0] * C + pos[..., 1]
    s, _ = lin.sort(dim=1)
    assert (s[:, 1:] != s[:, :-1]).all(), "vertex collision"
    glin = goals[..., 0] * C + goals[..., 1]
    gs, _ = glin.sort(dim=1)
    assert (gs[:, 1:] != gs[:, :-1]).all(), "duplicate goals"
    if lifelong:
        assert (glin != lin).all(), "lifelong goal under its own agent"
    lo = out.local_obs
    assert int(lo.max()) <= 4
    c = env.V // 2
    ok = (lo == 0) | (lo == 3) | (lo == 4)
    m = out.action_mask
    assert (m[..., 0] == 1).all()
    assert (m[..., 1].bool() == ok[..., c - 1, c]).all() and (m[..., 2].bool() == ok[..., c, c + 1]).all()
    assert (m[..., 3].bool() == ok[..., c + 1, c]).all() and (m[..., 4].bool() == ok[..., c, c - 1]).all()
    assert (lo[..., c, c] != 1).all() and (lo[..., c, c] != 2).all(), "own cell is never obstacle/other agent"
    # window obstacle cells agree with the map
    assert ((out.reward * 2).round() == out.reward * 2).all(), "rewards are multiples of 0.5"


def test_c3_full_size_invariants_and_determinism():
    """65 536 envs x 16 agents, 32x32, lifelong, auto-reset, fused masked sampler: invariants hold at
    every probe, and a second run from the same seed is bit-identical (no races)."""
    import torch

    runs = []
    for _ in range(2):
        env = make(c3_cfg(), 65536)
        out = env.reset()
        check_invariants(env, out, True)
        acts = env.sample_actions(masked=True)
        env.fuse_sampler("masked")
        arrivals = 0
        for s in range(150):
            out = env.step(acts, auto_reset=True)
            if s % 25 == 0 or s in (63, 64, 65):
                check_invariants(env, out, True)
                arrivals += int(out.info[:, nat.I_GOALS_REACHED_STEP].sum())
        assert env.poll_errors() == 0 and arrivals > 0
        w = env.state["env_words"]
        assert int(w[:, nat.W_EPISODES].min()) == 2 and int(w[:, nat.W_STEP_COUNT].max()) == 150 - 128
        m = env.metrics_vector().cpu().numpy()
        assert m[0] == 2 * 65536 and m[2] == 64 * m[0]  # episodes, length_sum
        runs.append({k: v.clone() for k, v in env.state.items()} | {"obs": out.local_obs.clone()})
        env.close()
    for k in runs[0]:
        assert torch.equal(runs[0][k], runs[1][k]), f"run-to-run difference in {k}"


def test_shard_invariance_global_env_ids():
    """Env e behaves identically whether it lives in one big batch or in a shard with env_id_base."""
    import torch

    cfg = c3_cfg(steps_per_episode=20)
    full = make(cfg, 96)
    shards = [make(cfg, 32, env_id_base=32 * r) for r in range(3)]
    envs = [full] + shards
    for e in envs:
        e.reset()
        e._a = e.sample_actions(masked=True)
        e.fuse_sampler("masked")
    for _ in range(45):
        for e in envs:
            e.step(e._a, auto_reset=True)
    for k in ("positions", "goals", "starts", "agent_flags", "lock_moved", "env_words"):
        cat = torch.cat([s.state[k] for s in shards])
        assert torch.equal(full.state[k], cat), k
    total = sum(s.metrics_vector().cpu().numpy() for s in shards)
    assert np.allclose(full.metrics_vector().cpu().numpy(), total)


def test_fused_sampler_equals_standalone_sampler():
    import torch

    cfg = c3_cfg(steps_per_episode=30)
    a, b = make(cfg, 300), make(cfg, 300)
    a.reset()
    b.reset()
    acts_a = a.sample_actions(masked=True)
    a.fuse_sampler("masked")
    for s in range(70):
        acts_b = b.sample_actions(masked=True)
        assert torch.equal(acts_a, acts_b), f"step {s}"
        oa = a.step(acts_a, auto_reset=True)
        ob = b.step(acts_b, auto_reset=True)
        ok = torch.gather(ob.action_mask.long(), 2, acts_b.long().unsqueeze(-1))
        assert torch.equal(oa.local_obs, ob.local_obs)
    assert torch.equal(a.state["positions"], b.state["positions"])
    # masked draws are always legal moves; unmasked draws are uniform over 0..4
    acts = b.sample_actions(masked=True)
    assert (torch.gather(b.out["action_mask"].long(), 2, acts.long().unsqueeze(-1)) == 1).all()
    big = make(cfg, 20000)
    big.reset()
    hist = torch.bincount(big.sample_actions(masked=False).flatten().long(), minlength=5).cpu().numpy()
    exp = 20000 * 16 / 5
    assert (((hist - exp) ** 2) / exp).sum() < 25, hist  # chi2, 4 dof


def test_philox_reset_layouts_are_uniform_and_distinct():
    """ENV:267-282 equivalence in distribution: 2N distinct free cells, every slot uniform over the
    free cells, start/goal slots pairwise distinct."""
    from dl_reference_models_b200 import maps

    grid = maps.get_grid("ReferenceModel-1-4")  # 13 free cells, N=4 -> 8 of 13 drawn
    env = make({"num_agents": 4, "sensor_range": 1, "seed": 11, "grid": grid}, 30000)
    env.reset()
    starts = env.state["starts"].cpu().numpy().astype(int)
    goals = env.state["goals"].cpu().numpy().astype(int)
    assert np.array_equal(starts, env.state["positions"].cpu().numpy())
    cells = np.concatenate([starts, goals], axis=1)
    lin = cells[..., 0] * 7 + cells[..., 1]
    assert (grid[cells[..., 0], cells[..., 1]] == 0).all()
    assert all(len(set(r)) == 8 for r in lin[:2000].tolist())
    s = np.sort(lin, axis=1)
    assert (s[:, 1:] != s[:, :-1]).all()
    free = np.flatnonzero(grid.reshape(-1) == 0)
    for slot in range(8):  # chi2 with 12 dof, 99.99% quantile ~ 39
        hist = np.array([(lin[:, slot] == f).sum() for f in free])
        exp = 30000 / 13
        assert (((hist - exp) ** 2) / exp).sum() < 45, (slot, hist)
    # joint of two slots: uniform over ordered distinct pairs (13*12 = 156 cells, 155 dof)
    pair = lin[:, 0] * 64 + lin[:, 5]
    _, counts = np.unique(pair, return_counts=True)
    assert len(counts) == 156
    exp = 30000 / 156
    assert (((counts - exp) ** 2) / exp).sum() < 240
    # a second reset draws a different layout (the per-env Philox counter advances)
    env.reset()
    assert (env.state["starts"].cpu().numpy() != starts).any(axis=(1, 2)).mean() > 0.95


def test_philox_goal_draw_is_uniform_over_candidates():
    """ENV:284-304: new goal uniform over free cells that are neither occupied nor a goal."""
    import torch

    grid = np.zeros((3, 3), np.uint8)
    cfg = {"num_agents": 2, "sensor_range": 1, "lifelong_mapf": True, "seed": 5, "grid": grid,
           "steps_per_episode": 1000}
    B = 24000
    env = make(cfg, B)
    starts = np.array([[0, 0], [2, 2]], np.int16)
    goals = np.array([[0, 1], [2, 0]], np.int16)
    env.reset(starts=starts, goals=goals)
    out = env.step(torch.tensor([[2, 0]], dtype=torch.int8).expand(B, 2).contiguous())  # agent 0 RIGHT -> arrives
    assert (out.step_flags & nat.SF_GOAL_REASSIGNED).bool().all()
    g = env.state["goals"].cpu().numpy().astype(int)
    assert (g[:, 1] == [2, 0]).all()
    lin = g[:, 0, 0] * 3 + g[:, 0, 1]
    # candidates: 9 cells - occupied {(0,1),(2,2)} - other goal {(2,0)} = 6 cells
    cand = sorted(set(range(9)) - {1, 8, 6})
    hist = np.array([(lin == c).sum() for c in cand])
    assert hist.sum() == B
    exp = B / 6
    assert (((hist - exp) ** 2) / exp).sum() < 30, hist  # chi2, 5 dof


def test_device_error_flags():
    import torch

    from dl_reference_models_b200 import maps

    env = make({"num_agents": 4, "sensor_range": 2, "seed": 1, "env_name": "ReferenceModel-2-1"}, 8)
    env.reset()
    bad = torch.zeros((8, 4), dtype=torch.int8)
    bad[3, 2] = 7
    env.step(bad)
    with pytest.raises(ValueError, match="Invalid action"):
        env.raise_on_device_errors()
    assert env.poll_errors() == 0  # polled flags are cleared
    with pytest.raises(nat.MapfError, match="free cells"):
        make({"num_agents": 6, "seed": 1, "grid": maps.get_grid("ReferenceModel-1-1")}, 4)  # ENV:270-275
    # no candidate cell for a lifelong reassignment (ENV:296-298)
    grid = np.zeros((1, 3), np.uint8)
    cfg = {"num_agents": 2, "sensor_range": 1, "lifelong_mapf": True, "deterministic": True, "grid": grid,
           "starts": [[0, 0], [0, 2]], "goals": [[0, 1], [0, 0]], "seed": 3}
    env = make(cfg, 2)
    env.step(torch.tensor([[2, 0], [2, 0]], dtype=torch.int8))
    with pytest.raises(RuntimeError, match="No valid cell"):
        env.raise_on_device_errors()


def test_wait_for_graph_cycles():
    """The pointer-jumping output (no reference counterpart): agent i waits for j iff i
    failed to move and j sits on i's intended cell; agents on a cycle of that graph are flagged."""
    import torch

    grid = np.zeros((4, 4), np.uint8)
    # agents 0..3 on a 2x2 block pushing clockwise (a rotation no sequential order can execute),
    # agent 4 pushes into agent 0 (waits on the cycle, not part of it), agent 5 idles
    starts = [[0, 0], [0, 1], [1, 1], [1, 0], [0, 0 + 2], [3, 3]]
    starts[4] = [2, 0]
    goals = [[3, 0], [3, 1], [3, 2], [2, 3], [0, 3], [1, 3]]
    cfg = {"num_agents": 6, "sensor_range": 1, "deterministic": True, "grid": grid, "starts": starts,
           "goals": goals, "seed": 1}
    env = make(cfg, 3)
    acts = torch.tensor([[2, 3, 4, 1, 1, 0]], dtype=torch.int8).expand(3, 6).contiguous()
    out = env.step(acts)
    asf = out.agent_step_flags.cpu().numpy()
    cyc = (asf & nat.ASF_WFG_CYCLE) != 0
    # sequential semantics: 0 blocked by 1, 1 by 2, 2 by 3 (not moved yet); 3 moves up? its target (0,0) is held by 0
    assert (env.state["positions"].cpu().numpy()[0] == np.array(starts)).all(), "nobody can move"
    assert cyc[0].tolist() == [True, True, True, True, False, False]
    assert ((asf & nat.ASF_FAILED_MOVE) != 0)[0].tolist() == [True] * 5 + [False]
    assert (out.step_flags.cpu().numpy() & nat.SF_WFG_CYCLE).all()
    assert int(out.info[0, nat.I_WFG_CYCLE_STEPS]) == 1
    # swap attempt = 2-cycle; a follower of a blocked chain is not on a cycle
    cfg2 = dict(cfg, num_agents=3, starts=[[0, 0], [0, 1], [0, 2]], goals=[[3, 3], [3, 2], [3, 1]])
    env2 = make(cfg2, 2)
    out2 = env2.step(torch.tensor([[2, 4, 4]], dtype=torch.int8).expand(2, 3).contiguous())
    cyc2 = (out2.agent_step_flags.cpu().numpy() & nat.ASF_WFG_CYCLE) != 0
    assert cyc2[0].tolist() == [True, True, False]
    # random play: every flagged agent failed to move, env flag == any agent flag
    env3 = make(c3_cfg(num_agents=32, steps_per_episode=50), 2000)
    env3.reset()
    a3 = env3.sample_actions(masked=False)
    env3.fuse_sampler("random")
    seen = 0
    for _ in range(60):
        o = env3.step(a3, auto_reset=True)
        f = o.agent_step_flags
        on_cycle = (f & nat.ASF_WFG_CYCLE) != 0
        assert ((f[on_cycle] & nat.ASF_FAILED_MOVE) != 0).all()
        assert torch.equal(on_cycle.any(dim=1), (o.step_flags & nat.SF_WFG_CYCLE) != 0)
        assert (on_cycle.sum(dim=1) != 1).all(), "a cycle has at least two members"
        seen += int(on_cycle.any(dim=1).sum())
    assert seen > 0


def test_state_snapshot_restore_and_observe():
    """get_state/set_state (checkpoint/resume of env state) reproduce the trajectory."""
    import torch

    cfg = c3_cfg(steps_per_episode=40)
    env = make(cfg, 128)
    env.reset()
    acts = [torch.randint(0, 5, (128, 16), dtype=torch.int8, generator=torch.Generator().manual_seed(s))
            for s in range(30)]
    for a in acts[:10]:
        env.step(a)
    snap = env.get_state()
    ref = [env.step(a).local_obs.clone() for a in acts[10:]]
    final = env.get_state()
    env.set_state(snap)
    again = [env.step(a).local_obs.clone() for a in acts[10:]]
    assert all(torch.equal(x, y) for x, y in zip(ref, again))
    for k, v in final.items():
        assert torch.equal(v, env.state[k]), k
