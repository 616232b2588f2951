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
