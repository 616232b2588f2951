"""CPU: the oracle against fixtures recorded from the live reference for states no legal step produces
(tests/golden/injected.npz, made by tests/golden/make_golden_injected.py): two / three agents injected onto one
cell -- the collision penalty of ENV:658-666 and the owner-grid semantics around it -- and the eight maps and
start / goal tables of get_grid.py against tests/golden/maps.npz."""
import numpy as np
import pytest

from dl_reference_models_b200 import maps
from oracle.oracle import OracleEnv
from trace_utils import GOLDEN

MAPS = ("ReferenceModel-1-1", "ReferenceModel-1-2", "ReferenceModel-1-3", "ReferenceModel-1-4", "ReferenceModel-2-1",
        "ReferenceModel-2-1-b", "ReferenceModel-2-2", "ReferenceModel-3-1")


def injected_config(lifelong: bool) -> dict:
    return {"num_agents": 6, "sensor_range": 2, "steps_per_episode": 5, "lifelong_mapf": lifelong,
            "deadlock_window_steps": 2, "livelock_window_steps": 3}


def load_injected():
    with np.load(GOLDEN / "injected.npz") as z:
        return {k: z[k] for k in z.files}


@pytest.mark.parametrize("lifelong", [False, True])
def test_oracle_replays_injected_colocation_traces(lifelong):
    g = load_injected()
    tag = "lifelong" if lifelong else "episodic"
    grid = g["grid"]
    S, T = g[f"{tag}_actions"].shape[:2]
    penalised = 0
    for sc in range(S):
        e = OracleEnv(injected_config(lifelong), grid)
        e.reset(1, g[f"{tag}_starts"][sc], g[f"{tag}_goals"][sc])
        e.set_state(positions=g[f"{tag}_injected_positions"][sc])
        for t in range(T):
            r = e.step(g[f"{tag}_actions"][sc, t], goal_rank=g[f"{tag}_goal_rank"][sc, t] if lifelong else None)
            st = e.state()
            ctx = f"{tag} scenario {sc} step {t}"
            assert np.array_equal(st["positions"], g[f"{tag}_positions"][sc, t]), ctx
            assert np.array_equal(st["goals"], g[f"{tag}_goals_after"][sc, t]), ctx
            assert np.array_equal(r.local_obs, g[f"{tag}_local_obs"][sc, t]), ctx
            assert np.array_equal(r.action_mask, g[f"{tag}_action_mask"][sc, t]), ctx
            assert np.abs(r.reward.astype(np.float64) - g[f"{tag}_reward"][sc, t]).max() <= 1e-6, ctx
            assert int(r.terminated[0]) == int(g[f"{tag}_terminated"][sc, t]), ctx
            assert int(r.truncated[0]) == int(g[f"{tag}_truncated"][sc, t]), ctx
            assert np.array_equal(r.moved, g[f"{tag}_moved"][sc, t]), ctx
            assert np.array_equal(r.failed_move, g[f"{tag}_failed_move"][sc, t]), ctx
            assert np.array_equal(r.blocking, g[f"{tag}_blocking"][sc, t]), ctx
            n = 14 if lifelong else 12
            assert np.array_equal(r.info_all[:n], g[f"{tag}_info_all"][sc, t, :n]), ctx
            penalised += int((r.reward <= -1.0).sum())
            if r.terminated[0] or r.truncated[0]:
                break
    assert penalised > 100


def test_maps_and_tables_match_the_reference_tables():
    """Every cell of all eight grids and every deterministic start / goal table (1..5 agents, including which
    counts the reference rejects) -- not just the free-cell counts."""
    with np.load(GOLDEN / "maps.npz") as z:
        ref = {k: z[k] for k in z.files}
    for name in MAPS:
        key = name.replace("ReferenceModel-", "m").replace("-", "_")
        assert np.array_equal(maps.get_grid(name), ref[f"{key}_grid"]), name
        assert maps.get_grid(name).dtype == np.uint8
        for n in range(1, 6):
            for what, fn in (("starts", maps.get_start_positions), ("goals", maps.get_goal_positions)):
                if f"{key}_{what}_{n}_error" in ref:
                    with pytest.raises(Exception) as ei:
                        fn(name, n)
                    assert type(ei.value).__name__ == str(ref[f"{key}_{what}_{n}_error"]), (name, what, n)
                else:
                    d = fn(name, n)
                    got = np.array([d[f"agent_{i}"] for i in range(n)], np.int16)
                    assert np.array_equal(got, ref[f"{key}_{what}_{n}"]), (name, what, n)
