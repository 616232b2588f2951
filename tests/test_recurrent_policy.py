"""SURVEY 8f N1 as the survey wrote it: the LSTM-64 policy of src/agents/ppo.py:67-75 (previous action and reward fed
to the cell, max_seq_len 32), the compact rollout batch (34 B per agent-step instead of 112 B of float32 features) and
the N-GPU learner's gradient all-reduce.  CPU tests use a stub env; the GPU test runs the real kernels."""
import os

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from dl_reference_models_b200.recurrent import (RecurrentActionMaskPolicy, RecurrentCollector, allreduce_gradients,
                                                features_from_channels, ppo_update_recurrent)
from dl_reference_models_b200.rollout import FLOAT_MIN


def test_policy_shape_masking_and_reference_hyperparameters():
    torch.manual_seed(0)
    pol = RecurrentActionMaskPolicy(28)   # fcnet_hiddens [64, 64], lstm_cell_size 64, prev action + prev reward
    assert pol.cell == 64 and pol.lstm.input_size == 64 + 5 + 1 and pol.lstm.hidden_size == 64
    M = 11
    f, mask = torch.randn(M, 28), torch.randint(0, 2, (M, 5), dtype=torch.int8)
    mask[:, 0] = 1
    lg, v, (h, c) = pol.step(f, mask, torch.zeros(M, dtype=torch.int64), torch.zeros(M), pol.initial_state(M))
    assert lg.shape == (M, 5) and v.shape == (M,) and h.shape == (M, 64) and c.shape == (M, 64)
    raw = pol.logits(h)
    assert torch.equal(lg, raw + torch.clamp(torch.log(mask.float() + 1e-6), min=FLOAT_MIN))   # action_mask_model.py:51-64
    assert float(torch.softmax(lg, -1)[mask == 0].max()) < 1e-4


def test_sequence_equals_steps_and_resets_cut_the_state():
    torch.manual_seed(1)
    pol = RecurrentActionMaskPolicy(12, hiddens=(16,), cell=8)
    T, M = 9, 5
    f, mask = torch.randn(T, M, 12), torch.ones(T, M, 5, dtype=torch.int8)
    pa, pr = torch.randint(0, 5, (T, M)), torch.randn(T, M)
    resets = torch.zeros(T, M, dtype=torch.bool)
    resets[0] = True
    resets[4, 2] = True
    lg, v, _ = pol.sequence(f, mask, pa, pr, resets, pol.initial_state(M))
    h, c = pol.initial_state(M)
    for t in range(T):
        keep = (~resets[t]).float().unsqueeze(-1)
        l2, v2, (h, c) = pol.step(f[t], mask[t], pa[t] * ~resets[t], pr[t] * ~resets[t], (h * keep, c * keep))
        assert torch.allclose(lg[t], l2, atol=1e-6) and torch.allclose(v[t], v2, atol=1e-6)
    # env 2 after its reset at t = 4 behaves like a fresh sequence: nothing before t = 4 leaks in
    lg2, _, _ = pol.sequence(f[4:, 2:3], mask[4:, 2:3], pa[4:, 2:3], pr[4:, 2:3], resets[4:, 2:3], pol.initial_state(1))
    assert torch.allclose(lg[4:, 2:3], lg2, atol=1e-6)


def test_features_from_channels_is_the_flat_observation_without_mask():
    lo = torch.randint(0, 5, (3, 2, 5, 5), dtype=torch.uint8)
    gd, bp = torch.randn(3, 2, 2), torch.randint(0, 2, (3, 2), dtype=torch.uint8)
    f = features_from_channels(lo, gd, bp)
    assert f.shape == (3, 2, 28) and f.dtype == torch.float32
    assert torch.equal(f[..., :25], lo.reshape(3, 2, 25).float()) and torch.equal(f[..., 25:27], gd)
    assert torch.equal(f[..., 27], bp.float())


class _StubEnv:
    """Enough of BatchedMapfEnv for the collector: random channels, episodes of 7 steps."""

    def __init__(self, B, N, seed):
        from dl_reference_models_b200.batched_env import StepOutput

        self.B, self.N, self.V, self.device, self._t, self._so = B, N, 5, torch.device("cpu"), 0, StepOutput
        self._g = torch.Generator().manual_seed(seed)

    def _make(self, done):
        B, N, g = self.B, self.N, self._g
        mask = torch.randint(0, 2, (B, N, 5), dtype=torch.int8, generator=g)
        mask[..., 0] = 1
        d = torch.full((B,), int(done), dtype=torch.uint8)
        return self._so(torch.randint(0, 5, (B, N, 5, 5), dtype=torch.uint8, generator=g), torch.randn(B, N, 2, generator=g),
                        torch.randint(0, 2, (B, N), dtype=torch.uint8, generator=g), mask, torch.randn(B, N, generator=g),
                        d, d.clone(), d.clone(), torch.zeros(B, N, dtype=torch.uint8), torch.zeros(B, 16, dtype=torch.int32))

    def _output(self):
        return self._make(False)

    def step(self, actions, auto_reset=True):
        self._t += 1
        return self._make(self._t % 7 == 0)


def test_compact_rollout_and_recurrent_ppo_update_on_a_stub_env():
    torch.manual_seed(3)
    env = _StubEnv(6, 4, 0)
    pol = RecurrentActionMaskPolicy(28, hiddens=(32,), cell=16)
    col = RecurrentCollector(env, pol, max_seq_len=8)
    batch = col.collect(16)
    assert batch.bytes_per_agent_step() == 25 + 8 + 1 + 5       # vs 4 * 28 + 5 for float32 feature rows
    assert batch.chunk_h.shape == (2, 24, 16) and int(batch.dones.sum()) == 2 * 6
    assert bool(batch.resets[0].all()) and bool(batch.resets[7].all()) and not bool(batch.resets[1].any())
    assert int(batch.prev_actions[7].abs().sum()) == 0            # zeroed behind a reset
    opt = torch.optim.Adam(pol.parameters(), lr=1e-3)
    before = [p.detach().clone() for p in pol.parameters()]
    stats = ppo_update_recurrent(pol, opt, batch, max_seq_len=8, epochs=2, minibatch=64)
    assert all(np.isfinite(v) for v in stats.values())
    assert any(not torch.equal(a, b) for a, b in zip(before, pol.parameters()))
    batch2 = col.collect(8)                                       # state carries over between collects
    assert not torch.equal(batch2.chunk_h[0], torch.zeros_like(batch2.chunk_h[0]))


def _ddp_worker(rank, world, port, q):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    pol = RecurrentActionMaskPolicy(6, hiddens=(8,), cell=4)
    torch.manual_seed(100 + rank)
    x = torch.randn(5, 6)
    lg, v, _ = pol.step(x, torch.ones(5, 5), torch.zeros(5, dtype=torch.int64), torch.zeros(5), pol.initial_state(5))
    (lg.sum() + v.sum()).backward()
    allreduce_gradients(pol, world)
    q.put((rank, [p.grad.detach().numpy().copy() for p in pol.parameters()]))   # plain arrays: no shared-memory handles
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_allreduce_two_ranks_gloo():
    """The N-GPU learner's collective on CPU: after allreduce_gradients every rank holds the mean of the per-rank
    gradients (= the gradient of the mean loss over both shards)."""
    import socket

    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    # reference: both shards in one process
    torch.manual_seed(0)
    pol = RecurrentActionMaskPolicy(6, hiddens=(8,), cell=4)
    grads = []
    for r in range(2):
        pol.zero_grad()
        torch.manual_seed(100 + r)
        x = torch.randn(5, 6)
        lg, v, _ = pol.step(x, torch.ones(5, 5), torch.zeros(5, dtype=torch.int64), torch.zeros(5), pol.initial_state(5))
        (lg.sum() + v.sum()).backward()
        grads.append([p.grad.clone() for p in pol.parameters()])
    for i in range(len(grads[0])):
        want = (grads[0][i] + grads[1][i]) / 2
        assert np.allclose(got[0][i], want.numpy(), atol=1e-6) and np.allclose(got[1][i], want.numpy(), atol=1e-6)


@pytest.mark.gpu
def test_recurrent_rollout_and_update_on_the_device():
    from dl_reference_models_b200 import maps
    from dl_reference_models_b200.batched_env import BatchedMapfEnv

    cfg = {"num_agents": 16, "sensor_range": 2, "steps_per_episode": 40, "lifelong_mapf": True, "seed": 4,
           "grid": maps.random_obstacle_grid(32, 32, 0.30, 2026, min_free=32)}
    env = BatchedMapfEnv(cfg, 256, "cuda:0")
    env.reset()
    torch.manual_seed(0)
    pol = RecurrentActionMaskPolicy(env.flat_obs_dim(include_action_mask=False)).to(env.device)
    col = RecurrentCollector(env, pol, max_seq_len=32)
    batch = col.collect(64)
    assert batch.local_obs.shape == (64, 256, 16, 5, 5) and batch.local_obs.is_cuda
    legal = torch.gather(batch.masks.long(), 3, batch.actions.unsqueeze(-1))
    assert float((legal == 0).float().mean()) < 1e-4
    assert int(batch.dones.sum()) == 256 and bool(batch.resets[40].all())
    assert env.poll_errors() == 0
    opt = torch.optim.Adam(pol.parameters(), lr=1e-3)
    stats = ppo_update_recurrent(pol, opt, batch, epochs=1, minibatch=1024, max_minibatches=6)
    assert all(np.isfinite(v) for v in stats.values())
