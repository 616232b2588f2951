"""Shortest-path distance table (bit-parallel BFS kernel) and per-agent goal path lengths against a plain Python BFS."""
from collections import deque

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def bfs_table(grid):
    R, C = grid.shape
    cells = R * C
    t = np.full((cells, cells), 255, np.uint8)
    for s in range(cells):
        sr, sc = divmod(s, C)
        if grid[sr, sc]:
            continue
        dist = {(sr, sc): 0}
        dq = deque([(sr, sc)])
        while dq:
            r, c = dq.popleft()
            for dr, dc in ((-1, 0), (1, 0), (0, -1), (0, 1)):
                rr, cc = r + dr, c + dc
                if 0 <= rr < R and 0 <= cc < C and not grid[rr, cc] and (rr, cc) not in dist:
                    dist[(rr, cc)] = dist[(r, c)] + 1
                    dq.append((rr, cc))
        for (r, c), d in dist.items():
            t[s, r * C + c] = min(d, 255)
    return t


@pytest.mark.parametrize("shape,density,seed", [((32, 32), 0.30, 2026), ((10, 20), 0.0, 0), ((7, 31), 0.35, 3), ((32, 5), 0.2, 4)])
def test_distance_table_and_goal_path_lengths(shape, density, seed):
    from dl_reference_models_b200 import maps
    from dl_reference_models_b200.batched_env import BatchedMapfEnv

    grid = maps.get_grid("ReferenceModel-2-1") if density == 0.0 else maps.random_obstacle_grid(*shape, density, seed, min_free=12)
    env = BatchedMapfEnv({"grid": grid, "num_agents": 6, "sensor_range": 1, "seed": seed}, 300)
    env.reset()
    ref = bfs_table(grid)
    got = env.distance_table().cpu().numpy()
    assert np.array_equal(got, ref)
    pos = env.state["positions"].cpu().numpy().astype(int)
    gl = env.state["goals"].cpu().numpy().astype(int)
    C = grid.shape[1]
    d = ref[gl[..., 0] * C + gl[..., 1], pos[..., 0] * C + pos[..., 1]].astype(np.int16)
    d[d == 255] = -1
    assert np.array_equal(env.goal_path_lengths().cpu().numpy(), d)
    # a lower bound indeed: greedy masked play never reaches a goal in fewer steps than the table says
    assert (d[d >= 0] >= np.abs(pos - gl).sum(-1)[d >= 0]).all()   # never below the Manhattan distance


MAPS = ("ReferenceModel-1-1", "ReferenceModel-1-2", "ReferenceModel-1-3", "ReferenceModel-1-4", "ReferenceModel-2-1",
        "ReferenceModel-2-1-b", "ReferenceModel-2-2", "ReferenceModel-3-1")


@pytest.mark.parametrize("name", MAPS)
def test_distance_table_against_the_reference_planners(name):
    """tests/golden/planners.npz (make_golden_planners.py): the reference's own space-time A* (scripts/cbs.py:22-137,
    no constraints) on 60 (start, goal) pairs of every reference map gives the distances the GPU table holds; the
    complete CBS solutions (scripts/cbs.py:240) for the deterministic tables cost every agent at least its table
    distance, and exactly it where the root node was conflict-free."""
    from pathlib import Path

    import torch

    from dl_reference_models_b200 import maps
    from dl_reference_models_b200.batched_env import BatchedMapfEnv

    with np.load(Path(__file__).resolve().parent / "golden" / "planners.npz") as z:
        fx = {k: z[k] for k in z.files}
    key = name.replace("ReferenceModel-", "m").replace("-", "_")
    grid = maps.get_grid(name)
    C = grid.shape[1]
    env = BatchedMapfEnv({"grid": grid, "num_agents": 2, "sensor_range": 1, "seed": 1}, 4)
    table = env.distance_table().cpu().numpy()
    pr = fx[f"{key}_pairs"].astype(int)
    got = table[pr[:, 0] * C + pr[:, 1], pr[:, 2] * C + pr[:, 3]].astype(np.int16)
    got[got == 255] = -1
    assert np.array_equal(got, fx[f"{key}_astar_len"])
    if f"{key}_cbs_starts" in fx:
        st, gl = fx[f"{key}_cbs_starts"].astype(int), fx[f"{key}_cbs_goals"].astype(int)
        n = len(st)
        env2 = BatchedMapfEnv({"grid": grid, "num_agents": n, "sensor_range": 1, "seed": 1}, 3)
        env2.reset(starts=torch.from_numpy(st.astype(np.int16)), goals=torch.from_numpy(gl.astype(np.int16)))
        d = env2.goal_path_lengths().cpu().numpy()
        assert (d == d[0]).all()
        assert np.array_equal(d[0], fx[f"{key}_cbs_root_len"])          # the root node's unconstrained paths
        if bool(fx[f"{key}_cbs_solved"]):
            assert (fx[f"{key}_cbs_path_len"] >= d[0]).all()             # lower bound of every agent's CBS path
            assert int(fx[f"{key}_cbs_makespan"]) >= int(d[0].max())    # ... and of the makespan
            if bool(fx[f"{key}_cbs_root_conflict_free"]):
                assert np.array_equal(fx[f"{key}_cbs_path_len"], d[0])
