"""Obstacle maps and deterministic start/goal tables for the MAPF environment.

The eight named maps and their start/goal tables are the *data* of the reference
(``src/environments/get_grid.py:17-727`` maps, ``:735-802`` starts, ``:805-872`` goals);
they are stored here as text rows (``#`` obstacle, ``.`` free) and decoded to ``uint8``
{0 free, 1 obstacle}.  Synthetic generators cover the BASELINE configs that have no
reference map (32x32 random obstacles, narrow corridors).
"""
from __future__ import annotations

import numpy as np

_GRID_ROWS = {
    "ReferenceModel-1-1": (
        "####.####",
        ".........",
    ),
    "ReferenceModel-1-2": (
        "..######..",
        "..........",
        "..######..",
    ),
    "ReferenceModel-1-3": (
        ".#.",
        ".#.",
        "...",
        ".#.",
        ".#.",
    ),
    "ReferenceModel-1-4": (
        "###.###",
        "###.###",
        "###.###",
        ".......",
        "###.###",
        "###.###",
        "###.###",
    ),
    "ReferenceModel-2-1": (
        "....................",
        "..#######..#######..",
        "..#######..#######..",
        "....................",
        "..#######..#######..",
        "..#######..#######..",
        "....................",
        "..#######..#######..",
        "..#######..#######..",
        "....................",
    ),
    "ReferenceModel-2-1-b": (
        "....................",
        "..##################",
        "..##################",
        "....................",
        "..##################",
        "..##################",
        "....................",
        "..##################",
        "..##################",
        "....................",
    ),
    "ReferenceModel-2-2": (
        "#.##.##.##.##.##.##.#",
        "#.##.##.##.##.##.##.#",
        "#.##.##.##.##.##.##.#",
        "#.##.##.##.##.##.##.#",
        "..##.##.##.##.##.##..",
        "..##.##.##.##.##.##..",
        "...#.##.##.##.##.#...",
        "#....##.##.##.##....#",
        ".....##.##.##.##.....",
        "###...#.##.##.#...###",
        "####....##.##....####",
        "........##.##........",
        "######...#.#...######",
        "#######.......#######",
        ".....................",
    ),
    "ReferenceModel-3-1": (
        "..............................",
        "..............................",
        "..###########################.",
        "..###########################.",
        "..###########################.",
        "..###########################.",
        "..............................",
        "................##############",
        "..############..##############",
        "..############..##############",
        "..############................",
        "..############..######.#####..",
        "..############..######.#####..",
        "..############..######.#####..",
        "..############..######.#####..",
        "..############..######.#####..",
        "..############..######.#####..",
        "..############..######.#####..",
        "..############..############..",
        "..............................",
    ),
}

_STARTS = {
    "ReferenceModel-1-1": ((1, 1), (1, 7)),
    "ReferenceModel-1-2": ((1, 1), (1, 8)),
    "ReferenceModel-1-3": ((1, 0), (1, 2)),
    "ReferenceModel-1-4": ((3, 1), (1, 3), (5, 3), (3, 5)),
    "ReferenceModel-2-1": ((5, 0), (3, 12), (6, 5), (6, 14)),
    "ReferenceModel-2-2": ((11, 1), (8, 13), (11, 5), (11, 14)),
    "ReferenceModel-3-1": ((15, 0), (6, 23), (16, 22), (16, 28)),
}

_GOALS = {
    "ReferenceModel-1-1": ((1, 8), (1, 0)),
    "ReferenceModel-1-2": ((1, 9), (1, 0)),
    "ReferenceModel-1-3": ((3, 2), (3, 0)),
    "ReferenceModel-1-4": ((3, 6), (6, 3), (0, 3), (3, 0)),
    "ReferenceModel-2-1": ((6, 6), (9, 3), (6, 0), (3, 3)),
    "ReferenceModel-2-2": ((14, 7), (14, 3), (11, 0), (4, 4)),
    "ReferenceModel-3-1": ((17, 22), (19, 4), (16, 0), (17, 0)),
}


def get_grid(env_name: str) -> np.ndarray:
    """uint8 [R, C] obstacle map; unknown names raise ValueError like get_grid.py:728-732."""
    try:
        rows = _GRID_ROWS[env_name]
    except KeyError as exc:
        msg = f"Unknown environment name: {env_name}"
        raise ValueError(msg) from exc
    return np.array([[1 if ch == "#" else 0 for ch in row] for row in rows], dtype=np.uint8)


def _table_lookup(table: dict, env_name: str, num_agents: int, what: str) -> dict:
    if env_name not in table:
        msg = f"Unknown environment name: {env_name}"
        raise ValueError(msg)
    entries = table[env_name]
    if num_agents > len(entries):
        msg = f"Requested number of agents ({num_agents}) exceeds available {what} in {env_name}"
        raise ValueError(msg)
    return {f"agent_{i}": entries[i] for i in range(num_agents)}


def get_start_positions(env_name: str, num_agents: int) -> dict:
    """Deterministic starts keyed by agent id (get_grid.py:735-802)."""
    return _table_lookup(_STARTS, env_name, num_agents, "positions")


def get_goal_positions(env_name: str, num_agents: int) -> dict:
    """Deterministic goals keyed by agent id (get_grid.py:805-872)."""
    return _table_lookup(_GOALS, env_name, num_agents, "goal positions")


def map_names() -> tuple:
    return tuple(_GRID_ROWS)


# ------------------------------------------------------------------ synthetic maps
def random_obstacle_grid(rows: int = 32, cols: int = 32, density: float = 0.30, seed: int = 2026,
                         min_free: int = 0) -> np.ndarray:
    """i.i.d. Bernoulli(density) obstacles; redrawn (seed+1, ...) until >= min_free free cells."""
    s = seed
    while True:
        rng = np.random.default_rng(s)
        grid = (rng.random((rows, cols)) < density).astype(np.uint8)
        if int((grid == 0).sum()) >= max(min_free, 1):
            return grid
        s += 1


def corridor_grid(rows: int = 32, cols: int = 32, connectors: tuple = (0, -1)) -> np.ndarray:
    """1-wide horizontal corridors on even rows, joined by full-height vertical connector
    columns (the 1-2 / 2-1-b motif): deadlock-heavy by construction."""
    grid = np.ones((rows, cols), dtype=np.uint8)
    grid[0::2, :] = 0
    for c in connectors:
        grid[:, c] = 0
    return grid
