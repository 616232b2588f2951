"""Rollout-loop CUDA kernels around the env step (BASELINE config 5): the action-mask MLP evaluated straight from the
env's output channels and sampled in one launch (``mapf_policy_act``: bf16 tensor-core MLP with f32 accumulation), and
GAE as a backwards scan (``mapf_gae``).  Thin ctypes wrappers; torch is the allocator / stream provider.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _native as nat


class FusedPolicy:
    """Inference view of an :class:`rollout.ActionMaskPolicy` (two 64-wide ReLU layers, 5 logits + value): call
    :meth:`refresh` after every optimizer step, :meth:`act` once per env step."""

    def __init__(self, policy, env, seed: int | None = None):
        self.policy, self.env = policy, env
        self._lib = nat.lib()
        self.F = int(env.flat_obs_dim(include_action_mask=False))
        self.with_bp = bool(env.include_blocking_pressure)
        assert self.F == env.V * env.V + 2 + (1 if self.with_bp else 0)
        lin = [m for m in policy.trunk if isinstance(m, torch.nn.Linear)]
        assert len(lin) == 2 and lin[0].out_features == 64 and lin[1].out_features == 64 and lin[0].in_features == self.F, \
            "the fused kernel implements the reference's 64-64 action-mask MLP"
        self._lin = lin
        n = int(self._lib.mapf_policy_weights_nbytes(self.F))
        if n < 0:
            raise nat.MapfError(n, self._lib.mapf_last_error().decode())
        self._host = np.zeros(n, np.uint8)
        self._dev = torch.zeros(n, dtype=torch.uint8, device=env.device)
        B, N = env.B, env.N
        dev = env.device
        self.actions = torch.zeros((B, N), dtype=torch.int8, device=dev)
        self.actions64 = torch.zeros((B, N), dtype=torch.int64, device=dev)
        self.logp = torch.zeros((B, N), dtype=torch.float32, device=dev)
        self.value = torch.zeros((B, N), dtype=torch.float32, device=dev)
        self.seed = int(env.cfg.seed if seed is None else seed) & (2 ** 64 - 1)
        self.counter = 0
        self.refresh()

    def refresh(self):
        """Re-pack the module's current float32 parameters (bf16, padded) and upload them."""
        f = lambda t: np.ascontiguousarray(t.detach().float().cpu().numpy())  # noqa: E731
        arrs = [f(self._lin[0].weight), f(self._lin[0].bias), f(self._lin[1].weight), f(self._lin[1].bias),
                f(self.policy.logits.weight), f(self.policy.logits.bias), f(self.policy.value.weight), f(self.policy.value.bias)]
        nat.check(self._lib.mapf_policy_pack_weights(self.F, *[a.ctypes.data_as(C.c_void_p) for a in arrs],
                                                     self._host.ctypes.data_as(C.c_void_p)))
        self._dev.copy_(torch.from_numpy(self._host))

    def act(self, out, logits_out: torch.Tensor | None = None, features_out: torch.Tensor | None = None,
            logp: torch.Tensor | None = None, value: torch.Tensor | None = None, actions64: torch.Tensor | None = None,
            mask_out: torch.Tensor | None = None):
        """One launch: features from ``out`` (the env's StepOutput) -> MLP -> masked categorical draw.
        Returns (actions int8 [B,N], logp, value); optional tensors receive the masked logits / the float feature
        block / a copy of the action masks (the rollout buffer's rows, so the loop needs no copy launches)."""
        env = self.env
        p = lambda t: None if t is None else C.c_void_p(t.data_ptr())  # noqa: E731
        logp = self.logp if logp is None else logp
        value = self.value if value is None else value
        actions64 = self.actions64 if actions64 is None else actions64
        self.counter += 1
        args = nat.MapfPolicyArgs(
            num_envs=env.B, num_agents=env.N, v2=env.V * env.V, feature_dim=self.F,
            no_masking=int(bool(self.policy.no_masking)), reserved=0, seed=self.seed, counter=self.counter,
            env_id_base=int(env.cfg.env_id_base), local_obs=p(out.local_obs), goal_delta=p(out.goal_delta),
            blocking_prev=p(out.blocking_prev) if self.with_bp else None, action_mask=p(out.action_mask),
            weights=p(self._dev), actions=p(self.actions), actions64=p(actions64), logp=p(logp), value=p(value),
            logits_out=p(logits_out), features_out=p(features_out), action_mask_out=p(mask_out))
        nat.check(self._lib.mapf_policy_act(C.byref(args), env._stream()))
        return self.actions, logp, value


def gae(rewards, values, dones, last_value, gamma: float = 0.99, lam: float = 0.95):
    """GAE over [T,B,N] float32 CUDA tensors (``dones`` [T,B] bool / uint8) with the scan kernel."""
    T, B, N = rewards.shape
    adv, ret = torch.empty_like(rewards), torch.empty_like(rewards)
    d = dones.to(torch.uint8).contiguous()
    st = C.c_void_p(torch.cuda.current_stream(rewards.device).cuda_stream)
    p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    nat.check(nat.lib().mapf_gae(p(rewards.contiguous()), p(values.contiguous()), p(d), p(last_value.contiguous()), p(adv),
                                 p(ret), int(T), int(B), int(N), C.c_float(gamma), C.c_float(lam), st))
    return adv, ret
