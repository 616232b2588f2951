"""mapf_step_host's bit-packed device->host transfer (csrc/mapf_pack_kernel.cuh + mapf_host_unpack.cpp):
the arrays delivered to the host buffers are bit for bit those of an unpacked device-side step, for every sensor
range, with and without goal-delta normalisation, with optional channels left out, for several thread / slice
counts; and fewer bytes cross PCIe."""
import ctypes as C

import pytest

from dl_reference_models_b200 import _native as nat

pytestmark = pytest.mark.gpu

OUT_KEYS = ("local_obs", "action_mask", "goal_delta", "blocking_prev", "reward", "terminated", "truncated",
            "step_flags", "agent_step_flags", "info")


def make(cfg, B):
    from dl_reference_models_b200.batched_env import BatchedMapfEnv

    return BatchedMapfEnv(cfg, B, "cuda:0")


def transfer_bytes(env):
    h2d, d2h = C.c_int64(0), C.c_int64(0)
    nat.check(nat.lib().mapf_host_transfer_bytes(env._h, C.byref(h2d), C.byref(d2h)))
    return h2d.value, d2h.value


def run_host_vs_device(cfg, B, steps, want=OUT_KEYS):
    import torch

    a, b = make(cfg, B), make(cfg, B)
    a.reset()
    b.reset()
    host = {k: torch.full_like(v, 111, device="cpu").pin_memory() for k, v in b.out.items()}
    cout = nat.MapfOutputs(**{k: (host[k].data_ptr() if k in want else None) for k in nat.OUTPUT_FIELDS})
    gen = torch.Generator().manual_seed(5)
    for s in range(steps):
        acts = torch.randint(0, 5, (B, a.N), dtype=torch.int8, generator=gen)
        oa = a.step(acts.cuda(), auto_reset=True)
        nat.check(nat.lib().mapf_step_host(b._h, C.c_void_p(acts.pin_memory().data_ptr()), None, None, C.byref(cout), 1))
        for k in want:
            assert torch.equal(getattr(oa, k).cpu(), host[k]), f"step {s}: {k}"
    for k in a.state:
        assert torch.equal(a.state[k], b.state[k]), f"state {k}"
    return a, b


def c3(**kw):
    from dl_reference_models_b200 import maps

    cfg = {"num_agents": 16, "sensor_range": 2, "steps_per_episode": 12, "lifelong_mapf": True, "seed": 4242,
           "grid": maps.random_obstacle_grid(32, 32, 0.30, 2026, min_free=32)}
    cfg.update(kw)
    return cfg


@pytest.mark.parametrize("sr", [1, 2, 3])
@pytest.mark.parametrize("normalize", [True, False])
def test_packed_transfer_is_exact(monkeypatch, sr, normalize):
    monkeypatch.setenv("MAPF_HOST_PACK", "1")  # auto would also pick it on a host with >= 12 cores per rank
    monkeypatch.setenv("MAPF_HOST_RAW_32NDS", "0")  # every env packed (the share is self-tuning otherwise)
    B = 8192 + 96
    cfg = c3(sensor_range=sr, normalize_goal_delta=normalize)
    a, b = run_host_vs_device(cfg, B, 20)
    v2 = (2 * sr + 1) ** 2
    rs = nat.lib().mapf_packed_record_bytes(v2)
    h2d, d2h = transfer_bytes(b)
    assert h2d == B * 16
    assert d2h == B * 16 * (rs + 2) + B * (3 + 64), "records + blocking_prev + agent_step_flags + per-env channels"
    assert rs < v2 + 5 + 8 + 4


def test_packed_transfer_with_channels_left_out_and_odd_shapes(monkeypatch):
    monkeypatch.setenv("MAPF_HOST_PACK", "1")  # auto would also pick it on a host with >= 12 cores per rank
    want = ("local_obs", "action_mask", "goal_delta", "reward", "terminated")
    for slices, threads in (("1", "1"), ("5", "3"), ("16", "7")):
        monkeypatch.setenv("MAPF_HOST_SLICES", slices)
        monkeypatch.setenv("MAPF_HOST_THREADS", threads)
        run_host_vs_device(c3(num_agents=7, lifelong_mapf=False), 8192 + 33, 14, want=want)
    # a map wider than the env-per-thread kernel takes (the packed path does not depend on the step kernel)
    from dl_reference_models_b200 import maps

    monkeypatch.delenv("MAPF_HOST_SLICES", raising=False)
    monkeypatch.delenv("MAPF_HOST_THREADS", raising=False)
    grid = maps.random_obstacle_grid(40, 100, 0.2, 3, min_free=300)
    run_host_vs_device({"num_agents": 12, "sensor_range": 2, "steps_per_episode": 9, "seed": 1, "grid": grid}, 8192, 12)


def test_plain_tail_share(monkeypatch):
    """Part of the batch may go as plain copies behind the records (balances PCIe against the host threads):
    any share delivers the same arrays."""
    B = 8192 + 160
    monkeypatch.setenv("MAPF_HOST_PACK", "1")
    for share in ("5", "16", None):
        if share is None:
            monkeypatch.delenv("MAPF_HOST_RAW_32NDS", raising=False)
        else:
            monkeypatch.setenv("MAPF_HOST_RAW_32NDS", share)
        _, b = run_host_vs_device(c3(), B, 40 if share is None else 10)
        rs = nat.lib().mapf_packed_record_bytes(25)
        d2h = transfer_bytes(b)[1]
        packed_all, plain_all = B * 16 * (rs + 2) + B * 67, B * 16 * 44 + B * 67
        if share == "16":
            tail = (B * 16 // 32) // 32 * 32
            assert d2h == packed_all + tail * 16 * (42 - rs)
        assert packed_all <= d2h <= plain_all


def test_unpacked_path_still_there(monkeypatch):
    """MAPF_HOST_PACK=0, a missing big channel, or a small batch: plain copies, same results."""
    B = 8192 + 64
    monkeypatch.setenv("MAPF_HOST_PACK", "0")
    _, b = run_host_vs_device(c3(), B, 6)
    assert transfer_bytes(b)[1] == B * 16 * 44 + B * 67
    monkeypatch.setenv("MAPF_HOST_PACK", "1")
    want = ("local_obs", "action_mask", "reward", "blocking_prev")
    _, b = run_host_vs_device(c3(), B, 6, want=want)
    assert transfer_bytes(b)[1] == B * 16 * (25 + 5 + 4 + 1)
    _, b = run_host_vs_device(c3(), 512, 6)
    assert transfer_bytes(b)[1] == 512 * 16 * 44 + 512 * 67


def test_python_step_host_api_matches_device_api_and_oracle(monkeypatch):
    """BatchedMapfEnv.step_host / reset_host (numpy in, numpy out): same results as the device-tensor API on a big
    batch (packed transfer), and as the CPU oracle on a small one (plain copies), from the same layout and actions."""
    import numpy as np
    import torch

    from oracle.oracle import OracleBatch

    monkeypatch.setenv("MAPF_HOST_PACK", "1")
    B = 8192 + 32
    cfg = c3(steps_per_episode=9)
    a, b = make(cfg, B), make(cfg, B)
    oa = a.reset()
    hb = b.reset_host()
    assert set(hb) == set(b.HOST_CHANNELS)
    for k in ("local_obs", "action_mask", "goal_delta", "blocking_prev"):
        assert np.array_equal(getattr(oa, k).cpu().numpy(), hb[k]), k
    rng = np.random.default_rng(0)
    for s in range(25):
        acts = rng.integers(0, 5, (B, a.N))
        oa = a.step(torch.as_tensor(acts, dtype=torch.int8), auto_reset=True)
        hb = b.step_host(acts)
        for k in b.HOST_CHANNELS:
            assert np.array_equal(getattr(oa, k).cpu().numpy(), hb[k]), f"step {s}: {k}"
    h2d, d2h = b.host_transfer_bytes()
    assert h2d == B * a.N and d2h < sum(v.nbytes for v in hb.values()) // 2

    # small batch against the CPU oracle (plain copies; the oracle draws the layout, no lifelong goal draws)
    from dl_reference_models_b200 import maps

    grid = maps.get_grid("ReferenceModel-2-1")
    cfg = {"num_agents": 4, "sensor_range": 2, "steps_per_episode": 40, "seed": 7, "grid": grid}
    Bs = 64
    ob = OracleBatch(cfg, grid, Bs, seed=5)
    env = make(cfg, Bs)
    ob.reset(2)
    st = ob.state()
    hb = env.reset_host(starts=st["starts"], goals=st["goals"])
    for k in ("local_obs", "action_mask", "goal_delta"):
        assert np.array_equal(ob.buf[k].reshape(hb[k].shape), hb[k]), f"reset: {k}"
    for s in range(40):
        acts = rng.integers(0, 5, (Bs, 4)).astype(np.int8)
        ob.step(acts)
        hb = env.step_host(acts, auto_reset=False)
        assert np.array_equal(ob.state()["positions"], env.state["positions"].cpu().numpy()), f"step {s}"
        for k in ("local_obs", "action_mask", "goal_delta", "terminated", "truncated"):
            assert np.array_equal(ob.buf[k].reshape(hb[k].shape), hb[k]), f"step {s}: {k}"
        assert np.allclose(ob.buf["reward"], hb["reward"], atol=1e-6), f"step {s}: reward"
        if (ob.buf["terminated"] | ob.buf["truncated"]).any():
            break


def test_pageable_host_arrays(monkeypatch):
    """Plain numpy arrays (not pinned) as destinations and as the action source: the packed path stages in the
    library's own pinned block and expands into whatever memory the caller gave."""
    import numpy as np
    import torch

    monkeypatch.setenv("MAPF_HOST_PACK", "1")
    B = 8192 + 64
    cfg = c3()
    a, b = make(cfg, B), make(cfg, B)
    a.reset()
    b.reset()
    host = {k: np.full(tuple(v.shape), 111, dtype=v.cpu().numpy().dtype) for k, v in b.out.items()}
    cout = nat.MapfOutputs(**{k: host[k].ctypes.data for k in nat.OUTPUT_FIELDS})
    rng = np.random.default_rng(3)
    for s in range(12):
        acts = rng.integers(0, 5, (B, a.N)).astype(np.int8)
        oa = a.step(torch.from_numpy(acts).cuda(), auto_reset=True)
        nat.check(nat.lib().mapf_step_host(b._h, C.c_void_p(acts.ctypes.data), None, None, C.byref(cout), 1))
        for k in OUT_KEYS:
            assert np.array_equal(getattr(oa, k).cpu().numpy(), host[k]), f"step {s}: {k}"


@pytest.mark.parametrize("sr,B", [(2, 8192 + 96), (1, 300), (3, 1000)])
def test_records_delivery_expands_to_the_plain_arrays(sr, B):
    """mapf_step_host_records: the four big channels arrive as bit-packed records (nothing expanded inside the call);
    mapf_unpack_records turns them into exactly the arrays a device-side step produced, the small channels arrive as
    plain arrays, and 13 B (sr 2) instead of 42 B per agent cross PCIe."""
    import numpy as np
    import torch

    cfg = c3(sensor_range=sr, steps_per_episode=9)
    a, b = make(cfg, B), make(cfg, B)
    a.reset()
    b.reset()
    gen = torch.Generator().manual_seed(8)
    for s in range(14):
        acts = torch.randint(0, 5, (B, a.N), dtype=torch.int8, generator=gen)
        oa = a.step(acts.cuda(), auto_reset=True)
        rec, small = b.step_host_records(acts, auto_reset=True)
        got = b.unpack_records(rec, threads=3)
        for k in ("local_obs", "action_mask", "goal_delta", "reward"):
            assert np.array_equal(getattr(oa, k).cpu().numpy(), got[k]), f"step {s}: {k}"
        for k in ("blocking_prev", "terminated", "truncated", "step_flags", "agent_step_flags", "info"):
            assert np.array_equal(getattr(oa, k).cpu().numpy(), small[k]), f"step {s}: {k}"
    for k in a.state:
        assert torch.equal(a.state[k], b.state[k]), f"state {k}"
    rs = nat.lib().mapf_packed_record_bytes((2 * sr + 1) ** 2)
    assert transfer_bytes(b)[1] == B * 16 * (rs + 2) + B * (3 + 64)


def test_non_temporal_expansion_and_measured_mode_are_exact(monkeypatch):
    """The packed delivery with the expansion going through cache-resident blocks and non-temporal stores
    (MAPF_HOST_NT=1), and the handle's own choice -- packed / packed + NT / plain, measured on its first sixteen calls --
    deliver the arrays of the device-side step, whatever mode a call happens to run in."""
    monkeypatch.setenv("MAPF_HOST_PACK", "1")
    monkeypatch.setenv("MAPF_HOST_NT", "1")
    _, b = run_host_vs_device(c3(), 8192 + 96, 12)
    assert nat.lib().mapf_host_transfer_mode(b._h) == 2
    monkeypatch.delenv("MAPF_HOST_PACK", raising=False)
    monkeypatch.delenv("MAPF_HOST_NT", raising=False)
    _, b = run_host_vs_device(c3(), 8192 + 96, 19)      # calls 0-15 calibrate, 16.. run in the mode that won
    assert nat.lib().mapf_host_transfer_mode(b._h) in (0, 1, 2)
    monkeypatch.setenv("MAPF_HOST_NT", "0")             # NT ruled out: two candidates, twelve calibration calls
    _, b = run_host_vs_device(c3(), 8192 + 96, 14)
    assert nat.lib().mapf_host_transfer_mode(b._h) in (0, 1)
