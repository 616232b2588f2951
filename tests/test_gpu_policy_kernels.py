"""Fused policy kernel (features -> bf16 tensor-core MLP -> masked categorical draw) and the GAE scan kernel against
plain PyTorch fp32 references of the same ops.  Tolerances: the kernel rounds inputs, weights and hidden activations
to bf16 and accumulates in f32; against a torch reference that rounds at the same points the masked logits and the
value agree to 2e-4 on average and 3e-2 at worst (a hidden activation next to a bf16 rounding boundary may flip by one
ulp with the accumulation order), against the pure fp32 module to 6e-2 abs (scaled with the logit magnitude)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def make_env(sr=2, B=2048, n=16):
    from dl_reference_models_b200 import maps
    from dl_reference_models_b200.batched_env import BatchedMapfEnv

    cfg = {"num_agents": n, "sensor_range": sr, "steps_per_episode": 64, "lifelong_mapf": True, "seed": 21,
           "grid": maps.random_obstacle_grid(32, 32, 0.3, 2026, min_free=2 * n)}
    env = BatchedMapfEnv(cfg, B, "cuda:0")
    out = env.reset()
    for _ in range(5):
        out = env.step(env.sample_actions(masked=True), auto_reset=True)
    return env, out


def bf16_reference(policy, feats, mask):
    import torch

    r = lambda t: t.to(torch.bfloat16).to(torch.float32)  # noqa: E731
    lin = [m for m in policy.trunk if isinstance(m, torch.nn.Linear)]
    h = r(feats)
    for l in lin:
        h = r(torch.relu(h @ r(l.weight).T + l.bias))
    logits = h @ r(policy.logits.weight).T + policy.logits.bias
    value = (h @ r(policy.value.weight).T + policy.value.bias).squeeze(-1)
    inf_mask = torch.where(mask != 0, torch.full_like(logits, 9.99999e-07), torch.full_like(logits, -13.815511))
    return logits + inf_mask, value


@pytest.mark.parametrize("sr", [1, 2, 3])
def test_fused_policy_matches_torch_forward(sr):
    import torch

    from dl_reference_models_b200.policy_kernels import FusedPolicy
    from dl_reference_models_b200.rollout import ActionMaskPolicy

    torch.manual_seed(0)
    env, out = make_env(sr=sr, B=1000 + 3)
    F = env.flat_obs_dim(include_action_mask=False)
    policy = ActionMaskPolicy(F).to(env.device)
    with torch.no_grad():
        for p in policy.parameters():
            p.mul_(3.0)   # larger logits: a sharper test of the arithmetic
    fused = FusedPolicy(policy, env)
    logits = torch.zeros((env.B, env.N, 5), device=env.device)
    feats = torch.zeros((env.B, env.N, F), device=env.device)
    acts, logp, value = fused.act(out, logits_out=logits, features_out=feats)
    ref_feats = env.flat_obs(include_action_mask=False)
    assert torch.equal(feats, ref_feats), "feature block"
    with torch.no_grad():
        ref_l, ref_v = bf16_reference(policy, ref_feats, out.action_mask)
        full_l, full_v = policy(ref_feats, out.action_mask)
    dl, dv = (logits - ref_l).abs(), (value - ref_v).abs()
    assert float(dl.max()) <= 3e-2 and float(dv.max()) <= 3e-2, (float(dl.max()), float(dv.max()))
    assert float(dl.mean()) <= 2e-4 and float(dv.mean()) <= 2e-4, (float(dl.mean()), float(dv.mean()))
    assert float((logits - full_l).abs().max()) <= 6e-2 * max(1.0, float(full_l.abs().max()) / 10)
    ref_logp = torch.log_softmax(logits, -1).gather(-1, acts.long().unsqueeze(-1)).squeeze(-1)
    assert float((logp - ref_logp).abs().max()) <= 1e-4
    assert torch.equal(fused.actions64, acts.long())
    assert (out.action_mask.gather(-1, acts.long().unsqueeze(-1)) == 1).float().mean() > 0.9999   # masked actions: p ~ 1e-6


def test_fused_policy_sampling_distribution():
    """Same observation in every env, different Philox streams: action frequencies follow softmax(logits)."""
    import torch

    from dl_reference_models_b200.policy_kernels import FusedPolicy
    from dl_reference_models_b200.rollout import ActionMaskPolicy

    torch.manual_seed(1)
    env, out = make_env(B=4096, n=4)
    F = env.flat_obs_dim(include_action_mask=False)
    policy = ActionMaskPolicy(F).to(env.device)
    fused = FusedPolicy(policy, env)
    for t in (out.local_obs, out.goal_delta, out.blocking_prev, out.action_mask):
        t.copy_(t[:1].expand_as(t).clone())
    logits = torch.zeros((env.B, env.N, 5), device=env.device)
    counts = torch.zeros((env.N, 5), device=env.device)
    draws = 0
    for _ in range(20):
        acts, _, _ = fused.act(out, logits_out=logits)
        counts += torch.nn.functional.one_hot(acts.long(), 5).sum(0)
        draws += env.B
    probs = torch.softmax(logits[0], -1)
    exp = probs * draws
    keep = exp > 5
    chi2 = (((counts - exp) ** 2 / exp.clamp_min(1e-9)) * keep).sum(-1)
    assert float(chi2.max()) < 40, (counts.tolist(), exp.tolist())   # <= 4 dof, 4 agents


def test_gae_kernel_matches_torch_scan():
    import torch

    from dl_reference_models_b200 import policy_kernels, rollout

    torch.manual_seed(2)
    T, B, N = 37, 513, 7
    dev = "cuda:0"
    rewards, values = torch.randn(T, B, N, device=dev), torch.randn(T, B, N, device=dev)
    dones = torch.rand(T, B, device=dev) < 0.1
    last = torch.randn(B, N, device=dev)
    adv, ret = policy_kernels.gae(rewards, values, dones, last, 0.99, 0.95)
    ref_adv, ref_ret = rollout.gae(rewards, values, dones, last, 0.99, 0.95)
    assert float((adv - ref_adv).abs().max()) <= 1e-4 and float((ret - ref_ret).abs().max()) <= 1e-4


def test_fused_collect_trains():
    import torch

    from dl_reference_models_b200.policy_kernels import FusedPolicy
    from dl_reference_models_b200.rollout import ActionMaskPolicy, collect_fused, ppo_update

    env, out = make_env(B=256, n=8)
    policy = ActionMaskPolicy(env.flat_obs_dim(include_action_mask=False)).to(env.device)
    fused = FusedPolicy(policy, env)
    batch = collect_fused(env, fused, 16, out)
    assert batch.features.shape == (16, 256, 8, 28) and torch.isfinite(batch.logp).all() and torch.isfinite(batch.values).all()
    opt = torch.optim.Adam(policy.parameters(), lr=1e-3)
    stats = ppo_update(policy, opt, batch, epochs=1, max_minibatches=4)
    assert all(np.isfinite(v) for v in stats.values())
    fused.refresh()
    fused.act(env._output())


def test_fused_collect_rows_are_what_a_manual_loop_sees():
    """collect_fused hands the rows of its [T,...] buffers to the two kernels (masks via the policy kernel, reward /
    terminated / truncated via the env step): they equal what a loop with explicit copies records, on an odd agent
    count (partial mask words at the end of the batch) and across auto-resets."""
    import torch

    from dl_reference_models_b200.policy_kernels import FusedPolicy
    from dl_reference_models_b200.rollout import ActionMaskPolicy, collect_fused

    T = 62
    ea, oa = make_env(B=203, n=7)
    eb, ob = make_env(B=203, n=7)
    torch.manual_seed(0)
    policy = ActionMaskPolicy(ea.flat_obs_dim(include_action_mask=False)).to(ea.device)
    fa, fb = FusedPolicy(policy, ea, seed=5), FusedPolicy(policy, eb, seed=5)
    batch = collect_fused(ea, fa, T, oa)
    out = ob
    for t in range(T):
        mask_t = out.action_mask.clone()
        a, lp, v = fb.act(out)
        assert torch.equal(batch.masks[t], mask_t), f"step {t}: masks"
        assert torch.equal(batch.actions[t], a.long()), f"step {t}: actions"
        assert torch.equal(batch.logp[t], lp), f"step {t}: logp"
        out = eb.step(a, auto_reset=True)
        assert torch.equal(batch.rewards[t], out.reward), f"step {t}: reward"
        assert torch.equal(batch.dones[t], (out.terminated | out.truncated).bool()), f"step {t}: dones"
    assert bool(batch.dones.any()), "the run crosses an episode end"
    for k in ea.state:
        assert torch.equal(ea.state[k], eb.state[k]), k
    with pytest.raises(ValueError):
        ea.step(None, reward_out=torch.zeros(3, device=ea.device))


def test_compact_fused_collector_equals_the_feature_storing_one():
    """FusedCollector(compact=True): the env step writes the observation channels of step t into row t + 1 of
    [T + 1, ...] buffers, the policy kernel reads row t and stores no float32 feature block (V^2 + 9 instead of
    4 (V^2 + 3) bytes per agent-step).  Same actions, log-probabilities, values, rewards, episode ends and env state as
    the feature-storing collector; the features expanded from the compact rows are its features; it trains; a second
    collect continues from the first."""
    import torch

    from dl_reference_models_b200.policy_kernels import FusedPolicy
    from dl_reference_models_b200.rollout import ActionMaskPolicy, CompactBatch, FusedCollector, ppo_update

    T = 40
    ea, _ = make_env(B=203, n=7)
    eb, _ = make_env(B=203, n=7)
    torch.manual_seed(0)
    policy = ActionMaskPolicy(ea.flat_obs_dim(include_action_mask=False)).to(ea.device)
    fa, fb = FusedPolicy(policy, ea, seed=5), FusedPolicy(policy, eb, seed=5)
    ca, cb = FusedCollector(ea, fa, T), FusedCollector(eb, fb, T, compact=True)
    for rnd in range(2):
        a, b = ca.collect(), cb.collect()
        assert isinstance(b, CompactBatch)
        for k in ("actions", "logp", "values", "rewards", "dones", "last_value", "masks"):
            assert torch.equal(getattr(a, k), getattr(b, k)), f"round {rnd}: {k}"
        assert torch.equal(a.features, b.features), f"round {rnd}: features"
        idx = torch.tensor([0, 5, 203 * 7 * 3 + 11, T * 203 * 7 - 1], device=ea.device)
        assert torch.equal(a.features.reshape(-1, a.features.shape[-1])[idx], b.features_of(idx))
        for k in ea.state:
            assert torch.equal(ea.state[k], eb.state[k]), k
        for k in ("local_obs", "goal_delta", "blocking_prev", "action_mask"):
            assert torch.equal(ea.out[k], eb.out[k]), f"round {rnd}: env buffer {k}"
    assert bool(b.dones.any())
    opt = torch.optim.Adam(policy.parameters(), lr=1e-3)
    stats = ppo_update(policy, opt, b, epochs=1, max_minibatches=3)
    assert all(np.isfinite(v) for v in stats.values())
