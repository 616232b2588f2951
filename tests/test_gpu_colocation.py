"""The co-location penalty (ENV:658-666): -1 to both members of every pair of agents that stand on one cell.
A legal step never produces such a state (ENV:516-521), so -- like the reference's own tests
(tests/test_reference_model_multi_agent_invariants.py:28-38) -- the state is injected: two and three agents on
one cell, into the oracle (oracle_set_state, owner grids rebuilt with last-index-wins, ENV:200-212) and into
both step kernels (BatchedMapfEnv.set_state).  Every channel is compared, including the observation of the
co-located agents themselves: the occupancy owner of a shared cell is its highest-index agent, so the others
see OTHER_AGENT (2) in the centre of their own window (ENV:737-739)."""
import numpy as np
import pytest

from gpu_utils import ORACLE_STEP_KEYS, STATE_KEYS, assert_batch_equal, gpu_channels

pytestmark = pytest.mark.gpu


def _inject(ob, env, positions):
    import torch

    for b, e in enumerate(ob.envs):
        e.set_state(positions=positions[b])
    st = env.get_state()
    st["positions"] = torch.from_numpy(positions).to(st["positions"].device)
    env.set_state(st)


@pytest.mark.parametrize("kind", ["lane", "env"])
@pytest.mark.parametrize("lifelong", [False, True])
def test_colocated_agents_pay_one_per_pair(kind, lifelong):
    import torch

    from dl_reference_models_b200 import maps
    from dl_reference_models_b200.batched_env import BatchedMapfEnv
    from oracle.oracle import OracleBatch

    grid = maps.get_grid("ReferenceModel-2-1")
    B, N = 96, 8
    cfg = {"num_agents": N, "sensor_range": 2, "steps_per_episode": 12, "lifelong_mapf": lifelong, "seed": 4,
           "deadlock_window_steps": 2, "livelock_window_steps": 4, "grid": grid, "step_kernel": kind}
    ob = OracleBatch(cfg, grid, B, seed=21)
    env = BatchedMapfEnv(cfg, B, "cuda:0")
    ob.reset(2)
    st = ob.state()
    env.reset(starts=st["starts"], goals=st["goals"])
    rng = np.random.default_rng(5)
    pos = st["positions"].copy()
    # env b: a pair (b even) or a triple (b odd) shares the cell of its lowest / middle / highest member
    groups = []
    for b in range(B):
        k = 2 if b % 2 == 0 else 3
        members = np.sort(rng.choice(N, size=k, replace=False))
        host = members[b % k]
        pos[b, members] = pos[b, host]
        groups.append(members)
    _inject(ob, env, pos)
    saw_penalty = 0
    for s in range(12):
        acts = rng.integers(0, 5, (B, N)).astype(np.int8)
        for b, members in enumerate(groups):
            acts[b, members] = 0      # the co-located agents stay put: the pair is still there after the step
        ob.step(acts)
        out = env.step(torch.from_numpy(acts), goal_rank=torch.from_numpy(ob.ranks.copy()) if lifelong else None)
        got = gpu_channels(env, out)
        ref = dict(ob.buf)
        ref.update({k: v for k, v in ob.state().items() if k != "blocking_prev"})
        assert_batch_equal(got, ref, ORACLE_STEP_KEYS + STATE_KEYS, f"step {s} ({kind})", lifelong, True)
        for b, members in enumerate(groups):
            k = len(members)
            r = ob.buf["reward"][b, members]
            # -1 per partner (ENV:664-665) on top of whatever else the step paid
            others = np.delete(np.arange(N), members)
            assert (r <= -(k - 1) + 1.5).all() and (ob.buf["reward"][b, others] >= -1.0).all()
            saw_penalty += int((r <= -(k - 1) + 0.5).sum())
        done = (ob.buf["terminated"] | ob.buf["truncated"]).astype(np.uint8)
        if done.any():
            break
    assert saw_penalty > B
    assert env.poll_errors() == 0


@pytest.mark.parametrize("kind", ["lane", "env"])
def test_known_answer_three_agents_one_cell(kind):
    """1-3 map, 3 agents injected onto (0,0), all NO_OP: rewards -2 each (two partners), nobody terminated."""
    import torch

    from dl_reference_models_b200 import maps
    from dl_reference_models_b200.batched_env import BatchedMapfEnv

    grid = maps.get_grid("ReferenceModel-1-4")
    free = np.argwhere(grid == 0).astype(np.int16)
    cfg = {"num_agents": 3, "sensor_range": 1, "steps_per_episode": 50, "grid": grid, "step_kernel": kind}
    env = BatchedMapfEnv(cfg, 4, "cuda:0")
    starts = np.broadcast_to(free[:3], (4, 3, 2)).copy()
    goals = np.broadcast_to(free[3:6], (4, 3, 2)).copy()
    env.reset(starts=starts, goals=goals)
    st = env.get_state()
    p = starts.copy()
    p[:, :, :] = free[0]
    st["positions"] = torch.from_numpy(p).cuda()
    env.set_state(st)
    out = env.step(torch.zeros((4, 3), dtype=torch.int8))
    assert out.reward.cpu().numpy().tolist() == [[-2.0, -2.0, -2.0]] * 4
    assert not out.terminated.any() and not out.truncated.any()
    obs = out.local_obs.cpu().numpy()
    # occupancy owner of the shared cell = agent 2 (last index wins, ENV:200-205): agents 0 and 1 see
    # OTHER_AGENT in their own centre, agent 2 does not
    assert (obs[:, 0, 1, 1] == 2).all() and (obs[:, 1, 1, 1] == 2).all() and (obs[:, 2, 1, 1] != 2).all()


@pytest.mark.parametrize("kind", ["lane", "env"])
@pytest.mark.parametrize("lifelong", [False, True])
def test_injected_colocation_traces_of_the_live_reference(kind, lifelong):
    """tests/golden/injected.npz (recorded from the Python reference, make_golden_injected.py): 48 scenarios x 4 steps
    as one batch.  The env-per-thread kernel keeps the reference's owner-grid semantics exactly (a bit of its
    occupancy board is "owner != -1"), so it replays every scenario; the lane-per-agent kernel resolves moves from
    agent lists and is exact as long as nobody enters or leaves a shared cell -- the "static" scenarios
    (DESIGN.md 9)."""
    import torch

    from dl_reference_models_b200.batched_env import BatchedMapfEnv
    from test_oracle_injected_golden import injected_config, load_injected

    g = load_injected()
    tag = "lifelong" if lifelong else "episodic"
    sel = np.arange(g[f"{tag}_actions"].shape[0]) if kind == "env" else np.flatnonzero(g[f"{tag}_static"])
    B, T = len(sel), g[f"{tag}_actions"].shape[1]
    cfg = dict(injected_config(lifelong), grid=g["grid"], step_kernel=kind)
    env = BatchedMapfEnv(cfg, B, "cuda:0")
    env.reset(starts=g[f"{tag}_starts"][sel], goals=g[f"{tag}_goals"][sel])
    st = env.get_state()
    st["positions"] = torch.from_numpy(g[f"{tag}_injected_positions"][sel]).cuda()
    env.set_state(st)
    alive = np.ones(B, bool)
    for t in range(T):
        out = env.step(torch.from_numpy(g[f"{tag}_actions"][sel, t]),
                       goal_rank=torch.from_numpy(g[f"{tag}_goal_rank"][sel, t]) if lifelong else None)
        got = gpu_channels(env, out, flat=False)
        n = 14 if lifelong else 12
        for k, ref in (("positions", "positions"), ("goals", "goals_after"), ("local_obs", "local_obs"),
                       ("action_mask", "action_mask"), ("moved", "moved"), ("failed_move", "failed_move"),
                       ("blocking", "blocking"), ("terminated", "terminated"), ("truncated", "truncated")):
            r = g[f"{tag}_{ref}"][sel, t]
            bad = np.flatnonzero([(not np.array_equal(got[k][b], r[b])) and alive[b] for b in range(B)])
            assert bad.size == 0, f"{kind} {tag} step {t}: {k} differs in scenarios {sel[bad]}\n got {got[k][bad[0]]}\n ref {r[bad[0]]}"
        assert np.abs(got["reward"].astype(np.float64) - g[f"{tag}_reward"][sel, t])[alive].max() <= 1e-6
        assert np.array_equal(got["info_all"][alive, :n], g[f"{tag}_info_all"][sel, t, :n][alive])
        alive &= ~((g[f"{tag}_terminated"][sel, t] | g[f"{tag}_truncated"][sel, t]).astype(bool))
        if not alive.any():
            break


@pytest.mark.parametrize("kind", ["lane", "env"])
def test_layout_override_with_a_repeated_cell_is_refused(kind):
    """A layout is 2N distinct cells (ENV:277 `replace=False`, the get_grid.py tables).  An override that names a goal
    (or a start) twice would need the goal-owner grid's last-index-wins rule (ENV:207-212), which the kernels do not
    keep: it fails loudly instead of silently showing OWN_GOAL where the reference shows OTHER_GOAL."""
    import torch

    from dl_reference_models_b200 import maps
    from dl_reference_models_b200.batched_env import BatchedMapfEnv

    grid = maps.get_grid("ReferenceModel-2-1")
    free = np.argwhere(grid == 0).astype(np.int16)
    env = BatchedMapfEnv({"num_agents": 4, "sensor_range": 2, "grid": grid, "step_kernel": kind}, 5, "cuda:0")
    starts, goals = np.broadcast_to(free[:4], (5, 4, 2)).copy(), np.broadcast_to(free[10:14], (5, 4, 2)).copy()
    env.reset(starts=starts, goals=goals)
    assert env.poll_errors() == 0
    bad = goals.copy()
    bad[3, 2] = bad[3, 0]                      # env 3: agents 0 and 2 share a goal
    env.reset(starts=starts, goals=bad)
    with pytest.raises(ValueError, match="repeats a cell"):
        env.raise_on_device_errors()
    bad = starts.copy()
    bad[1, 3] = bad[1, 1]                      # env 1: agents 1 and 3 share a start
    env.reset(starts=bad, goals=goals)
    with pytest.raises(ValueError, match="repeats a cell"):
        env.raise_on_device_errors()
    env.reset(starts=starts, goals=goals)
    env.step(torch.zeros((5, 4), dtype=torch.int8))
    assert env.poll_errors() == 0
