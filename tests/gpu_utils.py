"""Helpers of the `-m gpu` parity tests: turn BatchedMapfEnv tensors into the channel arrays the
oracle / golden traces use, and compare whole batches bit-exactly."""
from __future__ import annotations

import numpy as np

from dl_reference_models_b200 import _native as nat


def info_all_from_words(info: np.ndarray, num_agents: int) -> np.ndarray:
    """int32 [B,16] info words -> float64 [B,14] in the oracle's INFO_KEYS order (ENV:639-655)."""
    B = info.shape[0]
    out = np.zeros((B, 14), np.float64)
    out[:, :12] = info[:, :12]
    out[:, 12] = info[:, nat.I_COMPLETED_COUNT].astype(np.float64) / float(num_agents)
    out[:, 13] = info[:, nat.I_GOALS_REACHED_TOTAL].astype(np.float64) / np.maximum(info[:, nat.I_STEP_COUNT], 1)
    return out


def gpu_channels(env, out, flat=True) -> dict:
    """Everything comparable against the oracle, as numpy arrays with a leading B axis."""
    asf = out.agent_step_flags.cpu().numpy()
    af = env.state["agent_flags"].cpu().numpy()
    got = {
        "positions": env.state["positions"].cpu().numpy(),
        "goals": env.state["goals"].cpu().numpy(),
        "starts": env.state["starts"].cpu().numpy(),
        "local_obs": out.local_obs.cpu().numpy(),
        "action_mask": out.action_mask.cpu().numpy(),
        "goal_delta": out.goal_delta.cpu().numpy(),
        "blocking_prev": out.blocking_prev.cpu().numpy().astype(np.float32),
        "reward": out.reward.cpu().numpy(),
        "terminated": out.terminated.cpu().numpy(),
        "truncated": out.truncated.cpu().numpy(),
        "blocking": ((asf & nat.ASF_BLOCKING) != 0).astype(np.float32),
        "goal_reached_step": ((asf & nat.ASF_GOAL_REACHED) != 0).astype(np.float32),
        "moved": ((asf & nat.ASF_MOVED) != 0).astype(np.uint8),
        "failed_move": ((asf & nat.ASF_FAILED_MOVE) != 0).astype(np.uint8),
        "reached": ((af & nat.AF_REACHED) != 0).astype(np.uint8),
        "completed_once": ((af & nat.AF_COMPLETED_ONCE) != 0).astype(np.uint8),
        "info_all": info_all_from_words(out.info.cpu().numpy(), env.N),
        "step_flags": out.step_flags.cpu().numpy(),
        "agent_step_flags": asf,
    }
    if flat:
        got["flat_obs"] = env.flat_obs(True, True, True).cpu().numpy()
    return got


def one_env(got: dict, e: int) -> dict:
    return {k: v[e] for k, v in got.items()}


def first_mismatch(a: np.ndarray, b: np.ndarray):
    bad = np.argwhere(a != b)
    return None if bad.size == 0 else tuple(int(x) for x in bad[0])


def assert_batch_equal(got: dict, ref: dict, keys, ctx: str, lifelong: bool = True, lock: bool = True):
    """Bit-exact comparison of [B, ...] channel arrays (rewards: |diff| <= 1e-6)."""
    for k in keys:
        g, r = np.asarray(got[k]), np.asarray(ref[k])
        if k == "info_all":
            n = 14 if lifelong else 12
            g, r = g[:, :n], r[:, :n]
            if not lock:
                keep = [0, 1, 2, 3] + ([12, 13] if lifelong else [])
                g, r = g[:, keep], r[:, keep]
        if k == "reward":
            d = np.abs(g.astype(np.float64) - r.astype(np.float64))
            assert d.max() <= 1e-6, f"{ctx}: reward differs by {d.max()} at {np.unravel_index(d.argmax(), d.shape)}"
            continue
        assert g.shape == r.shape, f"{ctx}: {k} shape {g.shape} vs {r.shape}"
        assert g.dtype == r.dtype, f"{ctx}: {k} dtype {g.dtype} vs {r.dtype}"
        if not np.array_equal(g, r):
            idx = first_mismatch(g, r)
            e = idx[0]
            raise AssertionError(
                f"{ctx}: {k} mismatch in {int((g != r).sum())} elements, first at {idx}\n"
                f" got[{e}]=\n{g[e]}\n ref[{e}]=\n{r[e]}")


ORACLE_STEP_KEYS = ("local_obs", "action_mask", "goal_delta", "blocking_prev", "reward", "terminated",
                    "truncated", "blocking", "goal_reached_step", "moved", "failed_move", "info_all")
STATE_KEYS = ("positions", "goals", "reached", "completed_once")
