"""BASELINE config 5 (on-device rollout loop): the action-mask policy arithmetic and GAE on CPU,
the policy + CUDA env loop and one PPO update on the GPU."""
import numpy as np
import pytest
import torch

from dl_reference_models_b200.rollout import FLOAT_MIN, ActionMaskPolicy, gae


def test_masked_logits_follow_the_reference_formula():
    # reference: models/action_mask_model.py:51-64
    torch.manual_seed(0)
    pol = ActionMaskPolicy(28, hiddens=(64, 64))
    x = torch.randn(7, 3, 28)
    mask = torch.randint(0, 2, (7, 3, 5), dtype=torch.int8)
    mask[..., 0] = 1
    lg, v = pol(x, mask)
    z = pol.trunk(x)
    want = pol.logits(z) + torch.clamp(torch.log(mask.float() + 1e-6), min=FLOAT_MIN)
    assert torch.equal(lg, want) and v.shape == (7, 3)
    probs = torch.softmax(lg, -1)
    assert float(probs[mask == 0].max()) < 1e-5  # masked actions are (numerically) never sampled
    raw, _ = ActionMaskPolicy(28, no_masking=True)(x, mask)
    assert raw.shape == (7, 3, 5)


def test_gumbel_max_sampler_matches_softmax():
    from dl_reference_models_b200.rollout import sample_categorical

    torch.manual_seed(0)
    logits = torch.tensor([0.0, 1.0, -1.0, 2.0, -30.0]).expand(200000, 5)
    a, lp = sample_categorical(logits)
    freq = torch.bincount(a, minlength=5).float() / a.numel()
    assert torch.allclose(freq, torch.softmax(logits[0], 0), atol=5e-3)
    assert torch.allclose(lp, torch.log_softmax(logits, -1).gather(1, a[:, None])[:, 0])


def test_gae_matches_a_scalar_loop():
    rng = np.random.default_rng(1)
    T, B, N = 9, 4, 3
    r = torch.tensor(rng.normal(size=(T, B, N)), dtype=torch.float32)
    v = torch.tensor(rng.normal(size=(T, B, N)), dtype=torch.float32)
    d = torch.tensor(rng.random((T, B)) < 0.3)
    last = torch.tensor(rng.normal(size=(B, N)), dtype=torch.float32)
    adv, ret = gae(r, v, d, last, 0.99, 0.95)
    for b in range(B):
        for n in range(N):
            run, nxt = 0.0, float(last[b, n])
            for t in range(T - 1, -1, -1):
                nd = 0.0 if bool(d[t, b]) else 1.0
                delta = float(r[t, b, n]) + 0.99 * nxt * nd - float(v[t, b, n])
                run = delta + 0.99 * 0.95 * nd * run
                assert abs(run - float(adv[t, b, n])) < 1e-4
                nxt = float(v[t, b, n])
    assert torch.allclose(ret, adv + v)


@pytest.mark.gpu
def test_policy_env_loop_and_ppo_update_on_device():
    from dl_reference_models_b200 import maps
    from dl_reference_models_b200.batched_env import BatchedMapfEnv
    from dl_reference_models_b200.rollout import collect, ppo_update

    cfg = {"num_agents": 16, "sensor_range": 2, "steps_per_episode": 24, "lifelong_mapf": True, "seed": 4,
           "grid": maps.random_obstacle_grid(32, 32, 0.30, 2026, min_free=32)}
    env = BatchedMapfEnv(cfg, 512, "cuda:0")
    env.reset()
    torch.manual_seed(0)
    pol = ActionMaskPolicy(env.flat_obs_dim(include_action_mask=False)).to(env.device)
    batch = collect(env, pol, 40)
    assert batch.features.shape == (40, 512, 16, 28) and batch.features.is_cuda
    legal = torch.gather(batch.masks.long(), 3, batch.actions.unsqueeze(-1))
    # log(0 + 1e-6) = -13.8, not -inf (reference formula): illegal draws have probability ~1e-6 each
    assert float((legal == 0).float().mean()) < 1e-4, "the mask keeps the policy on legal moves"
    assert torch.isfinite(batch.logp).all() and torch.isfinite(batch.values).all()
    assert int(batch.dones.sum()) == 512  # one episode end (step 24) per env inside 40 steps
    assert env.poll_errors() == 0
    opt = torch.optim.Adam(pol.parameters(), lr=1e-3)
    before = [p.detach().clone() for p in pol.parameters()]
    stats = ppo_update(pol, opt, batch, epochs=1, minibatch=1024, max_minibatches=8)
    assert all(np.isfinite(v) for v in stats.values())
    assert any(not torch.equal(a, b) for a, b in zip(before, pol.parameters()))


def test_compact_batch_expands_features_from_padded_rows():
    """CompactBatch.features_of: float32 feature rows (window, goal delta, pressure) of arbitrary agent-steps out of
    uint8 / float32 rows that are padded apart (256-byte aligned rows), without ever reshaping the whole block."""
    import torch

    from dl_reference_models_b200.rollout import CompactBatch

    T, B, N, V = 3, 5, 7, 5
    gen = torch.Generator().manual_seed(0)

    def rows(shape, dtype):
        n = int(np.prod(shape))
        es = torch.empty((), dtype=dtype).element_size()
        stride = -(-(n * es) // 256) * 256 // es
        buf = torch.zeros((T + 1, stride), dtype=dtype)
        return buf[:, :n].view((T + 1,) + shape)

    lo, gd, bp = rows((B, N, V, V), torch.uint8), rows((B, N, 2), torch.float32), rows((B, N), torch.uint8)
    lo.copy_(torch.randint(0, 5, lo.shape, generator=gen).to(torch.uint8))
    gd.copy_(torch.rand(gd.shape, generator=gen))
    bp.copy_(torch.randint(0, 2, bp.shape, generator=gen).to(torch.uint8))
    assert not lo.is_contiguous()
    z = torch.zeros((T, B, N))
    b = CompactBatch(lo, gd, bp, torch.zeros((T, B, N, 5), dtype=torch.int8), z.long(), z, z, z, torch.zeros((T, B), dtype=torch.bool),
                     torch.zeros((B, N)))
    want = torch.cat([lo[:T].reshape(T, B, N, V * V).float(), gd[:T], bp[:T].float().unsqueeze(-1)], dim=-1)
    assert torch.equal(b.features, want)
    idx = torch.tensor([0, 1, B * N, T * B * N - 1, 17])
    assert torch.equal(b.features_of(idx), want.reshape(-1, V * V + 3)[idx])
    b2 = CompactBatch(lo, gd, None, b.masks, b.actions, z, z, z, b.dones, b.last_value)
    assert b2.features.shape[-1] == V * V + 2
