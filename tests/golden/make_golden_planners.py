"""Fixtures for the distance fields from the reference's own planners (build container only).

scripts/cbs.py cannot be imported here (matplotlib / pandas at module level), so its planner functions -- `a_star`
(:22-137, space-time A* of one agent under constraints), `detect_conflict` (:161-236), `cbs` (:240-...) and the
`CBSNode` class -- are lifted out of the UNMODIFIED source with `ast` and executed as they are.  Recorded per instance:
  * the length of the unconstrained single-agent A* path for many (start, goal) pairs on each of the eight maps
    (= the obstacle-aware shortest-path distance the GPU table must hold),
  * complete CBS solutions for the deterministic start / goal tables: per-agent path costs and the makespan (our
    per-agent distances are lower bounds of them, and equal where the root node had no conflict).

Run:  python tests/golden/make_golden_planners.py   ->  tests/golden/planners.npz
"""
from __future__ import annotations

import ast
import heapq
import sys
import time
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REPO = HERE.parents[1]
sys.dont_write_bytecode = True
sys.path.insert(0, str(REPO / "oracle" / "ref_stubs"))
sys.path.insert(1, "/root/reference")

from src.environments import get_grid as ref_get_grid  # noqa: E402

MAPS = ("ReferenceModel-1-1", "ReferenceModel-1-2", "ReferenceModel-1-3", "ReferenceModel-1-4", "ReferenceModel-2-1",
        "ReferenceModel-2-1-b", "ReferenceModel-2-2", "ReferenceModel-3-1")


def load_planner():
    src = Path("/root/reference/scripts/cbs.py").read_text()
    tree = ast.parse(src)
    keep = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and
            n.name in ("a_star", "CBSNode", "compute_cost", "detect_conflict", "cbs")]
    ns = {"heapq": heapq, "time": time, "MAX_CPU_TIME": 60, "start_time": time.process_time()}
    exec(compile(ast.Module(body=keep, type_ignores=[]), "/root/reference/scripts/cbs.py", "exec"), ns)
    return ns


def main():
    ns = load_planner()
    out = {}
    rng = np.random.default_rng(7)
    for name in MAPS:
        key = name.replace("ReferenceModel-", "m").replace("-", "_")
        grid = np.asarray(ref_get_grid.get_grid(name), np.uint8)
        g = grid.tolist()
        free = np.argwhere(grid == 0)
        pairs, lens = [], []
        for _ in range(60):
            a, b = free[rng.integers(len(free))], free[rng.integers(len(free))]
            path = ns["a_star"](g, (int(a[0]), int(a[1])), (int(b[0]), int(b[1])), [], 0)
            pairs.append([a[0], a[1], b[0], b[1]])
            lens.append(-1 if path is None else len(path) - 1)
        out[f"{key}_pairs"] = np.array(pairs, np.int16)
        out[f"{key}_astar_len"] = np.array(lens, np.int16)
        # CBS on the deterministic table of the map (2 agents on the 1-x maps, 4 otherwise; 2-1-b has none)
        n = 2 if name.startswith("ReferenceModel-1") else 4
        try:
            sp, gp = ref_get_grid.get_start_positions(name, n), ref_get_grid.get_goal_positions(name, n)
        except Exception:
            continue
        starts = [tuple(int(x) for x in sp[f"agent_{i}"]) for i in range(n)]
        goals = [tuple(int(x) for x in gp[f"agent_{i}"]) for i in range(n)]
        ns["start_time"] = time.process_time()
        sol = ns["cbs"](g, starts, goals)
        root = [ns["a_star"](g, starts[i], goals[i], [], i) for i in range(n)]
        out[f"{key}_cbs_starts"] = np.array(starts, np.int16)
        out[f"{key}_cbs_goals"] = np.array(goals, np.int16)
        out[f"{key}_cbs_root_len"] = np.array([len(p) - 1 for p in root], np.int16)
        out[f"{key}_cbs_solved"] = np.array(sol is not None)
        if sol is not None:
            out[f"{key}_cbs_path_len"] = np.array([len(p) - 1 for p in sol], np.int16)
            out[f"{key}_cbs_makespan"] = np.array(max(len(p) - 1 for p in sol), np.int16)
            out[f"{key}_cbs_root_conflict_free"] = np.array(ns["detect_conflict"](root) is None)
        print(name, "A* pairs", len(lens), "unreachable", int(sum(1 for x in lens if x < 0)),
              "| CBS", None if sol is None else [len(p) - 1 for p in sol])
    np.savez_compressed(HERE / "planners.npz", **out)


if __name__ == "__main__":
    main()
