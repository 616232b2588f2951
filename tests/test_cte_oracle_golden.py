"""The CTE oracle (oracle/cte_oracle.c, a C restatement of the reference's single-agent env view) replays the
traces recorded from the LIVE reference (tests/golden/cte_*.npz, tests/golden/make_golden_cte.py) bit-exactly:
flat float32 observation, float64 scalar reward, terminated / truncated, info, positions."""
from pathlib import Path

import numpy as np
import pytest

GOLDEN = sorted((Path(__file__).resolve().parent / "golden").glob("cte_*.npz"))


def load_cfg(z):
    cfg = {}
    for k, v in zip(z["cfg_keys"], z["cfg_vals"]):
        k, v = str(k), str(v)
        cfg[k] = v if k == "env_name" else (v == "True" if v in ("True", "False") else (float(v) if "." in v else int(v)))
    return cfg


@pytest.mark.parametrize("path", GOLDEN, ids=[p.stem for p in GOLDEN])
def test_cte_oracle_replays_reference_trace(path):
    from oracle.cte_oracle import CteOracleEnv

    z = np.load(path)
    cfg = load_cfg(z)
    env = CteOracleEnv(cfg, z["grid"])
    ep = 0
    obs = env.reset(z["reset_starts"][0], z["reset_goals"][0])
    assert np.array_equal(obs, z["reset_obs"][0])
    for t in range(len(z["actions"])):
        if z["reset_before"][t]:
            ep += 1
            obs = env.reset(z["reset_starts"][ep], z["reset_goals"][ep])
            assert np.array_equal(obs, z["reset_obs"][ep]), f"reset {ep}"
        obs, reward, term, trunc, info = env.step(z["actions"][t])
        assert obs.dtype == np.float32 and np.array_equal(obs, z["obs"][t]), f"obs, step {t}"
        assert reward == float(z["reward"][t]), f"reward, step {t}: {reward!r} vs {float(z['reward'][t])!r}"
        assert term == bool(z["terminated"][t]) and trunc == bool(z["truncated"][t]), f"done, step {t}"
        assert np.array_equal(info, z["info"][t]), f"info, step {t}"
        assert np.array_equal(env.positions, z["positions"][t]), f"positions, step {t}"
    assert ep == len(z["reset_obs"]) - 1
