"""CPU-side checks (no GPU): the C-ABI library builds for sm_100a, loads, exports every symbol
include/mapf_b200.h declares, fails loudly without a device; host-side logic (config translation,
spaces, shard ranges, metric means)."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

from dl_reference_models_b200 import _native as nat
from dl_reference_models_b200 import maps, metrics, spaces

REPO = Path(__file__).resolve().parents[1]
HEADER = (REPO / "include" / "mapf_b200.h").read_text()


def declared_functions():
    names = re.findall(r"^\s*(?:const\s+char\s*\*|int64_t|int)\s*\**\s*(mapf_[a-z_]+)\s*\(", HEADER, flags=re.M)
    return sorted(set(names))


def test_header_declares_what_python_binds():
    assert sorted(nat.EXPORTS) == declared_functions()


def test_library_exports_every_declared_symbol():
    L = nat.lib()
    for name in declared_functions():
        assert hasattr(L, name), name
    assert b"sm_100a" in L.mapf_version()
    out = subprocess.run(["nm", "-D", "--defined-only", str(nat.LIB_PATH)], capture_output=True, text=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    assert set(declared_functions()) <= exported


def test_library_contains_sm_100a_code_only():
    out = subprocess.run(["cuobjdump", "--list-elf", str(nat.LIB_PATH)], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, out


def test_struct_layouts_match_the_header():
    assert C.sizeof(nat.MapfConfig) == 16 * 4 + 8 + 8 + 4 + 4
    assert C.sizeof(nat.MapfState) == 10 * C.sizeof(C.c_void_p)
    assert C.sizeof(nat.MapfOutputs) == 10 * C.sizeof(C.c_void_p)
    for macro, val in (("MAPF_ENV_WORDS", nat.ENV_WORDS), ("MAPF_METRIC_COUNT", nat.METRIC_COUNT),
                       ("MAPF_INFO_WORDS", nat.INFO_WORDS)):
        assert macro in HEADER and val == 16
    # enum order of the info / env words the Python side indexes
    names = re.findall(r"^\s*MAPF_I_([A-Z_]+)\b", HEADER, flags=re.M)
    assert names.index("COMPLETED_COUNT") == nat.I_COMPLETED_COUNT and names.index("STEP_COUNT") == nat.I_STEP_COUNT
    names = re.findall(r"^\s*MAPF_W_([A-Z0-9_]+)\b", HEADER, flags=re.M)
    assert names.index("RNG_COUNTER") == nat.W_RNG_COUNTER and names.index("EPISODES") == nat.W_EPISODES
    names = re.findall(r"^\s*MAPF_M_([A-Z0-9_]+)\b", HEADER, flags=re.M)
    assert [n.lower() for n in names] == list(nat.METRIC_NAMES)


def _no_gpu():
    import torch

    return not torch.cuda.is_available()


@pytest.mark.skipif(not _no_gpu(), reason="checks the no-device failure path")
def test_no_device_fails_loudly_no_cpu_fallback():
    L = nat.lib()
    cfg = nat.make_config({"num_agents": 4, "seed": 1}, 10, 20, 8)
    h = C.c_void_p()
    rc = L.mapf_create(C.byref(cfg), C.byref(h))
    assert rc == nat.ERR_CUDA and not h.value
    assert b"no CPU fallback" in L.mapf_last_error()
    with pytest.raises(nat.MapfError):
        nat.check(rc)
    from dl_reference_models_b200.batched_env import BatchedMapfEnv
    from dl_reference_models_b200.reference_model import ReferenceModel

    with pytest.raises((RuntimeError, ValueError)):
        BatchedMapfEnv({"env_name": "ReferenceModel-2-1", "num_agents": 4}, 8, "cuda:0")
    with pytest.raises(ValueError):
        BatchedMapfEnv({"env_name": "ReferenceModel-2-1", "num_agents": 4}, 8, "cpu")
    with pytest.raises(nat.MapfError):
        ReferenceModel({"env_name": "ReferenceModel-2-1", "num_agents": 4})


def test_argument_validation_happens_before_any_device_work():
    L = nat.lib()
    h = C.c_void_p()
    for bad in ({"num_agents": 33}, {"num_agents": 4, "sensor_range": 4}, {"num_agents": 4, "sensor_range": 0},
                {"num_agents": 4, "deadlock_window_steps": 40}):
        cfg = nat.make_config(bad, 10, 20, 8)
        assert L.mapf_create(C.byref(cfg), C.byref(h)) == nat.ERR_UNSUPPORTED, bad
    cfg = nat.make_config({"num_agents": 4}, 300, 20, 8)
    assert L.mapf_create(C.byref(cfg), C.byref(h)) == nat.ERR_UNSUPPORTED
    cfg = nat.make_config({"num_agents": 4}, 10, 20, 0)
    assert L.mapf_create(C.byref(cfg), C.byref(h)) == nat.ERR_INVALID_ARG
    assert L.mapf_create(None, C.byref(h)) == nat.ERR_INVALID_ARG
    assert L.mapf_step(None, None, None, None, None, 0, None) == nat.ERR_INVALID_ARG
    assert L.mapf_launch_count(None) == 0 and L.mapf_destroy(None) == 0


def test_make_config_uses_the_reference_defaults():
    c = nat.make_config({}, 10, 20, 3)
    assert (c.num_agents, c.sensor_range, c.steps_per_episode) == (2, 1, 100)            # ENV:38-40
    assert (c.lifelong_mapf, c.enable_lock_metrics, c.deterministic, c.normalize_goal_delta) == (0, 1, 0, 1)
    assert (c.deadlock_window_steps, c.livelock_window_steps, c.lock_nearby_manhattan,
            c.lock_min_neighbors, c.lock_progress_epsilon_floor) == (8, 16, 2, 1, 1)    # ENV:56-60
    c = nat.make_config({"deadlock_window_steps": 0, "lock_min_neighbors": -3, "lock_progress_epsilon": 2.7,
                         "seed": 5}, 4, 4, 1)
    assert (c.deadlock_window_steps, c.lock_min_neighbors, c.lock_progress_epsilon_floor, c.seed) == (1, 1, 2, 5)
    assert nat.make_config({"lock_progress_epsilon": -0.5}, 4, 4, 1).lock_progress_epsilon_floor == -1


def test_maps_and_tables():
    free = {"ReferenceModel-1-1": 10, "ReferenceModel-1-2": 18, "ReferenceModel-1-3": 11, "ReferenceModel-1-4": 13,
            "ReferenceModel-2-1": 116, "ReferenceModel-2-1-b": 92, "ReferenceModel-2-2": 151, "ReferenceModel-3-1": 229}
    for name, f in free.items():
        g = maps.get_grid(name)
        assert g.dtype == np.uint8 and int((g == 0).sum()) == f, name
    with pytest.raises(ValueError, match="Unknown environment name"):
        maps.get_grid("x")
    with pytest.raises(ValueError, match="exceeds"):
        maps.get_start_positions("ReferenceModel-1-1", 3)
    s, g = maps.get_start_positions("ReferenceModel-2-1", 4), maps.get_goal_positions("ReferenceModel-2-1", 4)
    grid = maps.get_grid("ReferenceModel-2-1")
    assert all(grid[p] == 0 for p in list(s.values()) + list(g.values()))
    c = maps.corridor_grid(32, 32)
    assert c.shape == (32, 32) and (c[1::2, 1:-1] == 1).all() and (c[0::2] == 0).all()
    r = maps.random_obstacle_grid(32, 32, 0.3, 2026, min_free=32)
    assert r.shape == (32, 32) and 600 < int((r == 0).sum()) < 800


def test_fallback_spaces_contract():
    b = spaces.Box(low=0, high=4, shape=(5, 5), dtype=np.uint8)
    assert b.shape == (5, 5) and b.dtype == np.uint8 and b.contains(np.zeros((5, 5), np.uint8))
    assert not b.contains(np.zeros((5, 4), np.uint8)) and not b.contains(np.full((5, 5), 9, np.uint8))
    f = spaces.Box(low=np.zeros(3, np.float32), high=np.ones(3, np.float32), dtype=np.float32)
    assert f.contains(np.array([0, .5, 1], np.float32)) and not f.contains(np.array([0, .5, 1], np.float64))
    assert spaces.Discrete(5).n == 5 and spaces.Discrete(5).contains(4) and not spaces.Discrete(5).contains(5)
    mb = spaces.MultiBinary(5)
    assert mb.shape == (5,) and mb.dtype == np.int8


def test_shard_ranges_partition_the_envs():
    for total in (1, 7, 64, 65536, 65537):
        for world in (1, 2, 3, 8):
            spans = [metrics.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_metric_means():
    v = np.zeros(16)
    v[0], v[1], v[3], v[4], v[10] = 4, 10.0, 1, 6, 0.5
    m = metrics.summarize(v)
    assert m["episodes"] == 4 and m["return_mean"] == 2.5 and m["success_rate"] == 0.25
    assert m["goals_reached_mean"] == 1.5 and m["throughput_mean"] == 0.125
    assert metrics.summarize(np.zeros(16))["return_mean"] == 0.0
    assert not any(k.startswith("reserved") for k in m)


def _pack_block_numpy(obs, mask, drow, dcol, reward2):
    """The packed block of csrc/mapf_pack_kernel.cuh restated in numpy (the spec the host expansion is checked
    against): [n x bits][n x (int8 d_row, int8 d_col)][n x int8 2*reward]; bits = 3 per window cell, 8 cells ->
    3 bytes, the last cell shares its byte with the 5 mask bits."""
    n, v2 = obs.shape
    nfull = v2 // 8
    assert v2 % 8 == 1
    bits = np.zeros((n, nfull * 3 + 1), dtype=np.uint8)
    for c in range(nfull):
        w = np.zeros(n, dtype=np.uint32)
        for i in range(8):
            w |= (obs[:, c * 8 + i].astype(np.uint32) & 7) << (3 * i)
        for b in range(3):
            bits[:, 3 * c + b] = (w >> (8 * b)) & 0xFF
    w = obs[:, v2 - 1].astype(np.uint32) & 7
    for i in range(5):
        w |= (mask[:, i] != 0).astype(np.uint32) << (3 + i)
    bits[:, 3 * nfull] = w
    diff = np.stack([drow.astype(np.int8), dcol.astype(np.int8)], axis=1).view(np.uint8)
    return np.concatenate([bits.reshape(-1), diff.reshape(-1), reward2.astype(np.int8).view(np.uint8)])


@pytest.mark.parametrize("isa", ["generic", "bmi2", "vbmi", "vbmi+nt"])
@pytest.mark.parametrize("v2,threads,n", [(9, 1, 10007), (25, 1, 10007), (25, 5, 70001), (49, 3, 10007), (25, 2, 3), (9, 4, 64)])
def test_host_expansion_of_packed_block(monkeypatch, isa, v2, threads, n):
    """Every instruction-set variant of the expansion (capped by MAPF_HOST_ISA; a CPU without the extension falls
    back to the next one) reproduces the arrays exactly -- also the staged one with non-temporal whole-line stores
    (MAPF_HOST_NT=1, for hosts bound by their memory system), unaligned heads / tails and guard rows included."""
    monkeypatch.setenv("MAPF_HOST_ISA", isa.split("+")[0])
    monkeypatch.setenv("MAPF_HOST_NT", "1" if isa.endswith("+nt") else "0")
    L = nat.lib()
    rng = np.random.default_rng(v2 * 10 + threads)
    obs = rng.integers(0, 5, (n, v2), dtype=np.uint8)
    mask = rng.integers(0, 2, (n, 5), dtype=np.int8)
    drow = rng.integers(-128, 128, n)
    dcol = rng.integers(-128, 128, n)
    reward2 = rng.integers(-64, 4, n)
    packed = _pack_block_numpy(obs, mask, drow, dcol, reward2)
    assert packed.size == n * L.mapf_packed_record_bytes(v2)
    for den_row, den_col in ((31.0, 17.0), (1.0, 1.0)):
        o = np.full((n + 1, v2), 255, np.uint8)   # one guard row behind every array: nothing may spill
        m = np.full((n + 1, 5), 77, np.int8)
        gd = np.full((n + 1, 2), -7.0, np.float32)
        rw = np.full(n + 1, -7.0, np.float32)
        ptr = lambda a: a.ctypes.data_as(C.c_void_p)
        rc = L.mapf_unpack_records(ptr(packed), n, v2, threads, ptr(o), ptr(m), ptr(gd), ptr(rw), den_row, den_col)
        assert rc == 0
        assert np.array_equal(o[:n], obs) and np.array_equal(m[:n], mask)
        assert np.array_equal(gd[:n, 0], drow.astype(np.float32) / np.float32(den_row))
        assert np.array_equal(gd[:n, 1], dcol.astype(np.float32) / np.float32(den_col))
        assert np.array_equal(rw[:n], 0.5 * reward2.astype(np.float32))
        assert (o[n] == 255).all() and (m[n] == 77).all() and (gd[n] == -7.0).all() and rw[n] == -7.0
