"""The recurrent policy of the reference's PPO configuration and a multi-GPU learner step around the batched env
(plain PyTorch by design -- the policy is the user's model, the env transition inside the loop is the
CUDA kernel).

Reference: ``src/agents/ppo.py:67-75`` (``fcnet_hiddens [64, 64]``, ``use_lstm``, ``lstm_cell_size 64``,
``lstm_use_prev_action``, ``lstm_use_prev_reward``, ``max_seq_len 32``, ``vf_share_layers``) and ``:104-117``
(gamma 0.99, lambda 0.95, clip 0.05, lr 1e-3, entropy 1e-3, vf 0.5, minibatch 1024, 12 epochs).

* :class:`RecurrentActionMaskPolicy` -- MLP trunk -> LSTM(64) fed with the trunk output, the one-hot previous action
  and the previous reward -> logits (masked like ``models/action_mask_model.py:51-64``) and value heads.
* :class:`CompactRollout` / :func:`collect_recurrent` -- T env steps on the device; the batch keeps the env's
  *channels* (uint8 window 25 B + float32 goal delta 8 B + pressure flag 1 B = 34 B per agent-step at sensor range 2)
  instead of float32 feature rows (112 B), and expands them per minibatch (:func:`features_from_channels`), plus the
  LSTM state at the start of every ``max_seq_len`` chunk (truncated back-propagation through time, as RLlib does).
* :func:`ppo_update_recurrent` -- PPO on sequence minibatches; with ``world_size > 1`` every rank holds its own env
  shard and the gradients are all-reduced (mean) before each optimizer step: an N-GPU data-parallel learner whose
  only traffic is the ~50 k parameters' gradients (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
from torch import nn

from .rollout import FLOAT_MIN, gae, sample_categorical


def features_from_channels(local_obs: torch.Tensor, goal_delta: torch.Tensor, blocking_prev: torch.Tensor | None) -> torch.Tensor:
    """ENV:306-328 without the mask: [..., V, V] uint8, [..., 2] float32, [...] uint8 -> [..., V*V + 2 (+ 1)] float32."""
    parts = [local_obs.flatten(-2).float(), goal_delta.float()]
    if blocking_prev is not None:
        parts.append(blocking_prev.float().unsqueeze(-1))
    return torch.cat(parts, dim=-1)


class RecurrentActionMaskPolicy(nn.Module):
    def __init__(self, feature_dim: int, num_actions: int = 5, hiddens=(64, 64), cell: int = 64,
                 use_prev_action: bool = True, use_prev_reward: bool = True):
        super().__init__()
        layers, last = [], int(feature_dim)
        for h in hiddens:
            layers += [nn.Linear(last, int(h)), nn.ReLU()]
            last = int(h)
        self.trunk = nn.Sequential(*layers)
        self.num_actions, self.cell = int(num_actions), int(cell)
        self.use_prev_action, self.use_prev_reward = bool(use_prev_action), bool(use_prev_reward)
        self.lstm = nn.LSTMCell(last + (num_actions if use_prev_action else 0) + (1 if use_prev_reward else 0), cell)
        self.logits = nn.Linear(cell, num_actions)
        self.value = nn.Linear(cell, 1)   # vf_share_layers: the value head reads the same LSTM output

    def initial_state(self, n: int, device=None):
        z = torch.zeros((n, self.cell), device=device)
        return z, z.clone()

    def step(self, features, action_mask, prev_action, prev_reward, state):
        """One time step for a flat batch [M, ...]: returns (masked logits [M, A], value [M], new state)."""
        x = [self.trunk(features.float())]
        if self.use_prev_action:
            x.append(torch.nn.functional.one_hot(prev_action.long(), self.num_actions).to(x[0].dtype))
        if self.use_prev_reward:
            x.append(prev_reward.to(x[0].dtype).unsqueeze(-1))
        h, c = self.lstm(torch.cat(x, dim=-1), state)
        logits = self.logits(h) + torch.clamp(torch.log(action_mask.to(h.dtype) + 1e-6), min=FLOAT_MIN)
        return logits, self.value(h).squeeze(-1), (h, c)

    def sequence(self, features, action_mask, prev_action, prev_reward, resets, state):
        """[T, M, ...] inputs; ``resets`` [T, M] bool = the step starts a new episode (state, previous action and
        previous reward are zeroed in front of it).  Returns logits [T, M, A], values [T, M], final state."""
        T = features.shape[0]
        lg, vs = [], []
        h, c = state
        for t in range(T):
            keep = (~resets[t]).to(h.dtype).unsqueeze(-1)
            h, c = h * keep, c * keep
            pa = prev_action[t] * (~resets[t]).to(prev_action.dtype)
            pr = prev_reward[t] * (~resets[t]).to(prev_reward.dtype)
            l, v, (h, c) = self.step(features[t], action_mask[t], pa, pr, (h, c))
            lg.append(l)
            vs.append(v)
        return torch.stack(lg), torch.stack(vs), (h, c)


@dataclass
class CompactRollout:
    local_obs: torch.Tensor      # [T,B,N,V,V] uint8
    goal_delta: torch.Tensor     # [T,B,N,2]  float32
    blocking_prev: torch.Tensor  # [T,B,N]    uint8
    masks: torch.Tensor          # [T,B,N,5]  int8
    actions: torch.Tensor        # [T,B,N]    int64
    prev_actions: torch.Tensor   # [T,B,N]    int64 (action of the step before, 0 behind a reset)
    prev_rewards: torch.Tensor   # [T,B,N]    float32
    resets: torch.Tensor         # [T,B]      bool: step t is the first of an episode
    logp: torch.Tensor           # [T,B,N]
    values: torch.Tensor         # [T,B,N]
    rewards: torch.Tensor        # [T,B,N]
    dones: torch.Tensor          # [T,B]      bool: the episode ended AT step t (the env auto-reset)
    last_value: torch.Tensor     # [B,N]
    chunk_h: torch.Tensor        # [T/L, B*N, cell] LSTM state at the start of every max_seq_len chunk
    chunk_c: torch.Tensor

    def bytes_per_agent_step(self) -> int:
        per = lambda t: t[0, 0, 0].numel() * t.element_size()  # noqa: E731
        return per(self.local_obs) + per(self.goal_delta) + per(self.blocking_prev) + per(self.masks)


class RecurrentCollector:
    """Carries env outputs, LSTM state, previous action / reward across :meth:`collect` calls."""

    def __init__(self, env, policy: RecurrentActionMaskPolicy, max_seq_len: int = 32):
        self.env, self.policy, self.L = env, policy, int(max_seq_len)
        M = env.B * env.N
        self.state = policy.initial_state(M, env.device)
        self.prev_action = torch.zeros((env.B, env.N), dtype=torch.int64, device=env.device)
        self.prev_reward = torch.zeros((env.B, env.N), device=env.device)
        self.reset_next = torch.ones((env.B,), dtype=torch.bool, device=env.device)
        self.out = env._output()

    @torch.no_grad()
    def collect(self, steps: int) -> CompactRollout:
        env, pol, L = self.env, self.policy, self.L
        if steps % L:
            raise ValueError(f"steps ({steps}) must be a multiple of max_seq_len ({L})")
        if getattr(env, "_fused", 0):
            raise RuntimeError("turn the env's fused uniform sampler off (fuse_sampler(None)) before a policy rollout")
        B, N, V, dev, T = env.B, env.N, env.V, env.device, int(steps)
        z = lambda shape, dt: torch.empty(shape, dtype=dt, device=dev)  # noqa: E731
        r = CompactRollout(z((T, B, N, V, V), torch.uint8), z((T, B, N, 2), torch.float32), z((T, B, N), torch.uint8),
                           z((T, B, N, 5), torch.int8), z((T, B, N), torch.int64), z((T, B, N), torch.int64),
                           z((T, B, N), torch.float32), z((T, B), torch.bool), z((T, B, N), torch.float32),
                           z((T, B, N), torch.float32), z((T, B, N), torch.float32), z((T, B), torch.bool),
                           z((B, N), torch.float32), z((T // L, B * N, pol.cell), torch.float32),
                           z((T // L, B * N, pol.cell), torch.float32))
        out = self.out
        h, c = self.state
        for t in range(T):
            rs = self.reset_next
            keep = (~rs).repeat_interleave(N).to(h.dtype).unsqueeze(-1)
            h, c = h * keep, c * keep
            if t % L == 0:
                r.chunk_h[t // L], r.chunk_c[t // L] = h, c
            pa = self.prev_action * (~rs).unsqueeze(-1)
            pr = self.prev_reward * (~rs).unsqueeze(-1)
            r.local_obs[t], r.goal_delta[t], r.blocking_prev[t], r.masks[t] = out.local_obs, out.goal_delta, out.blocking_prev, out.action_mask
            r.prev_actions[t], r.prev_rewards[t], r.resets[t] = pa, pr, rs
            feats = features_from_channels(out.local_obs, out.goal_delta, out.blocking_prev).reshape(B * N, -1)
            lg, v, (h, c) = pol.step(feats, out.action_mask.reshape(B * N, 5), pa.reshape(-1), pr.reshape(-1), (h, c))
            a, lp = sample_categorical(lg)
            r.actions[t], r.logp[t], r.values[t] = a.reshape(B, N), lp.reshape(B, N), v.reshape(B, N)
            out = env.step(a.reshape(B, N).to(torch.int8), auto_reset=True)
            r.rewards[t] = out.reward
            done = (out.terminated | out.truncated).bool()
            r.dones[t] = done
            self.prev_action, self.prev_reward, self.reset_next = a.reshape(B, N), out.reward.clone(), done
        # bootstrap value of the observation after the last step (state is not advanced)
        rs = self.reset_next
        keep = (~rs).repeat_interleave(N).to(h.dtype).unsqueeze(-1)
        feats = features_from_channels(out.local_obs, out.goal_delta, out.blocking_prev).reshape(B * N, -1)
        _, lv, _ = pol.step(feats, out.action_mask.reshape(B * N, 5), (self.prev_action * (~rs).unsqueeze(-1)).reshape(-1),
                            (self.prev_reward * (~rs).unsqueeze(-1)).reshape(-1), (h * keep, c * keep))
        r.last_value.copy_(lv.reshape(B, N))
        self.state, self.out = (h, c), out
        return r


def collect_recurrent(env, policy, steps: int, max_seq_len: int = 32) -> CompactRollout:
    return RecurrentCollector(env, policy, max_seq_len).collect(steps)


def allreduce_gradients(module: nn.Module, world_size: int, group=None):
    """Mean of the gradients over the ranks, one flat all-reduce (the learner's only collective)."""
    if world_size <= 1:
        return
    import torch.distributed as dist

    grads = [p.grad for p in module.parameters() if p.grad is not None]
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat /= world_size
    at = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[at:at + n].view_as(g))
        at += n


def ppo_update_recurrent(policy: RecurrentActionMaskPolicy, optimizer, batch: CompactRollout, *, max_seq_len: int = 32,
                         clip: float = 0.05, vf_coeff: float = 0.5, entropy_coeff: float = 0.001, epochs: int = 12,
                         minibatch: int = 1024, gamma: float = 0.99, lam: float = 0.95, world_size: int = 1, group=None,
                         max_minibatches: int | None = None, generator: torch.Generator | None = None) -> dict:
    """PPO over sequence minibatches: ``minibatch`` time steps = ``minibatch / max_seq_len`` sequences of one agent,
    each re-run through the LSTM from its stored start state (truncated BPTT).  Every rank draws its own minibatches
    from its own shard; gradients are averaged over the ranks before each optimizer step."""
    L = int(max_seq_len)
    T, B, N = batch.actions.shape
    adv, ret = gae(batch.rewards, batch.values, batch.dones, batch.last_value, gamma, lam)
    adv = (adv - adv.mean()) / (adv.std() + 1e-8)
    nchunks, M = T // L, B * N
    per_mb = max(1, minibatch // L)

    def seq_view(x):   # [T,B,N,...] -> [nchunks, L, M, ...]
        return x.reshape(nchunks, L, M, *x.shape[3:])

    lo, gd, bp, mk = seq_view(batch.local_obs), seq_view(batch.goal_delta), seq_view(batch.blocking_prev), seq_view(batch.masks)
    ac, pa, pr = seq_view(batch.actions), seq_view(batch.prev_actions), seq_view(batch.prev_rewards)
    olp, av, rt = seq_view(batch.logp), seq_view(adv), seq_view(ret)
    rs = batch.resets.unsqueeze(-1).expand(T, B, N).reshape(nchunks, L, M)
    stats, done_mb = {}, 0
    total = nchunks * M
    for _ in range(epochs):
        perm = torch.randperm(total, device=ac.device, generator=generator) if generator is None or generator.device == ac.device \
            else torch.randperm(total, generator=generator).to(ac.device)
        for i in range(0, total, per_mb):
            idx = perm[i:i + per_mb]
            ck, m = idx // M, idx % M
            g = lambda x: x[ck, :, m].transpose(0, 1)   # noqa: E731  -> [L, S, ...]
            feats = features_from_channels(g(lo), g(gd), g(bp))   # expanded here, per minibatch: 34 B -> 112 B per row
            lg, v, _ = policy.sequence(feats, g(mk), g(pa), g(pr), g(rs), (batch.chunk_h[ck, m], batch.chunk_c[ck, m]))
            dist_ = torch.distributions.Categorical(logits=lg)
            lp = dist_.log_prob(g(ac))
            ratio = torch.exp(lp - g(olp))
            a_ = g(av)
            surr = torch.min(ratio * a_, torch.clamp(ratio, 1 - clip, 1 + clip) * a_)
            vf = (v - g(rt)).pow(2).mean()
            ent = dist_.entropy().mean()
            loss = -surr.mean() + vf_coeff * vf - entropy_coeff * ent
            optimizer.zero_grad(set_to_none=True)
            loss.backward()
            allreduce_gradients(policy, world_size, group)
            optimizer.step()
            stats = {"loss": float(loss.detach()), "vf_loss": float(vf.detach()), "entropy": float(ent.detach()),
                     "policy_loss": float(-surr.mean().detach())}
            done_mb += 1
            if max_minibatches is not None and done_mb >= max_minibatches:
                return stats
    return stats
