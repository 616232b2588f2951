"""The oracle's device-RNG replay (oracle_set_philox): Philox4x32-10 against the Random123 known-answer
vectors, and the draw discipline's basic properties (distinct layouts, uniformity, counter bookkeeping).
CPU only; the GPU side of the same draws is tests/test_gpu_benchmarked_mode_replay.py."""
import numpy as np

from dl_reference_models_b200 import maps
from oracle import oracle as orc


def test_philox4x32_10_known_answers():
    # Random123 kat_vectors, philox4x32 10 rounds
    assert [hex(x) for x in orc.philox_raw(0, 0, [0, 0, 0, 0])] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    assert [hex(x) for x in orc.philox_raw(0xFFFFFFFF, 0xFFFFFFFF, [0xFFFFFFFF] * 4)] == [
        "0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    assert [hex(x) for x in orc.philox_raw(0xA4093822, 0x299F31D0, [0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344])] == [
        "0xd16cfe09", "0x94fdcceb", "0x5001e420", "0x24126ea1"]


def test_philox_layouts_are_distinct_and_env_specific():
    grid = maps.get_grid("ReferenceModel-2-1")
    cfg = {"num_agents": 8, "sensor_range": 2, "lifelong_mapf": True, "seed": 5}
    ob = orc.OracleBatch(cfg, grid, 300)
    ob.set_philox(5, env_id_base=1000)
    ob.reset(2)
    st = ob.state()
    cells = np.concatenate([st["starts"], st["goals"]], axis=1).astype(int)
    lin = cells[..., 0] * grid.shape[1] + cells[..., 1]
    assert all(len(set(r)) == 16 for r in lin.tolist())                    # ENV:277 replace=False
    assert (grid[cells[..., 0], cells[..., 1]] == 0).all()
    assert len({tuple(r) for r in lin.tolist()}) > 290                     # keyed by the global env id
    c0 = ob.philox_counters()
    assert (c0 >= 1).all()                                                 # at least one round per draw
    ob2 = orc.OracleBatch(cfg, grid, 10)
    ob2.set_philox(5, env_id_base=1005)                                    # shard invariance: env 1005.. again
    ob2.reset(2)
    assert np.array_equal(ob2.state()["starts"], st["starts"][5:15])


def test_philox_masked_sampler_respects_mask_and_is_uniform():
    grid = maps.get_grid("ReferenceModel-2-1")
    cfg = {"num_agents": 4, "sensor_range": 2, "seed": 3}
    ob = orc.OracleBatch(cfg, grid, 64)
    ob.set_philox(3)
    ob.reset(2)
    hist = np.zeros(5)
    for c in range(1, 200):
        a = ob.sample_actions(c, masked=True)
        assert np.take_along_axis(ob.buf["action_mask"], a[..., None].astype(np.int64), 2).all()
        u = ob.sample_actions(c, masked=False)
        hist += np.bincount(u.ravel(), minlength=5)
    p = hist / hist.sum()
    assert np.abs(p - 0.2).max() < 0.01
