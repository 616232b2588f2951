"""Observation / action space descriptors of the MAPF env.

When ``gymnasium`` is importable (an RLlib install) its space classes are used, so RLlib sees the
real thing.  Otherwise these minimal stand-ins provide what the env contract needs
(``shape, dtype, low, high, n, contains, sample``): the reference builds its spaces at
``src/environments/reference_model_multi_agent.py:138-175,214-265``.
"""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - depends on the installation
    from gymnasium.spaces import Box, Discrete, MultiBinary  # type: ignore

    HAVE_GYMNASIUM = True
except Exception:  # gymnasium absent: self-contained descriptors
    HAVE_GYMNASIUM = False

    class _Space:
        shape: tuple = ()
        dtype = None

        def __contains__(self, x):
            return self.contains(x)

    class Box(_Space):
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.dtype = np.dtype(dtype)
            if shape is None:
                shape = np.broadcast_shapes(np.shape(low), np.shape(high))
            self.shape = tuple(int(s) for s in shape)
            self.low = np.broadcast_to(np.asarray(low), self.shape).astype(self.dtype)
            self.high = np.broadcast_to(np.asarray(high), self.shape).astype(self.dtype)

        def contains(self, x) -> bool:
            x = np.asarray(x)
            return bool(np.can_cast(x.dtype, self.dtype) and x.shape == self.shape
                        and np.all(x >= self.low) and np.all(x <= self.high))

        def sample(self, rng=None):
            rng = rng or np.random.default_rng()
            return rng.uniform(self.low, self.high).astype(self.dtype)

        def __repr__(self):
            return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"

    class Discrete(_Space):
        def __init__(self, n):
            self.n = int(n)
            self.shape = ()
            self.dtype = np.dtype(np.int64)

        def contains(self, x) -> bool:
            try:
                return 0 <= int(x) < self.n
            except (TypeError, ValueError):
                return False

        def sample(self, rng=None):
            rng = rng or np.random.default_rng()
            return int(rng.integers(self.n))

        def __repr__(self):
            return f"Discrete({self.n})"

    class MultiBinary(_Space):
        def __init__(self, n):
            self.n = int(n)
            self.shape = (int(n),)
            self.dtype = np.dtype(np.int8)

        def contains(self, x) -> bool:
            x = np.asarray(x)
            return bool(x.shape == self.shape and np.all((x == 0) | (x == 1)))

        def sample(self, rng=None):
            rng = rng or np.random.default_rng()
            return rng.integers(0, 2, self.shape).astype(self.dtype)

        def __repr__(self):
            return f"MultiBinary({self.n})"

try:  # pragma: no cover - depends on the installation
    from gymnasium.spaces import MultiDiscrete  # type: ignore
except Exception:
    class MultiDiscrete:  # joint action space of the single-agent (CTE) view
        def __init__(self, nvec):
            self.nvec = np.asarray(nvec, dtype=np.int64)
            self.shape = self.nvec.shape
            self.dtype = np.dtype(np.int64)

        def contains(self, x) -> bool:
            x = np.asarray(x)
            return bool(x.shape == self.shape and np.all(x >= 0) and np.all(x < self.nvec))

        def __contains__(self, x):
            return self.contains(x)

        def sample(self, rng=None):
            rng = rng or np.random.default_rng()
            return rng.integers(0, self.nvec).astype(self.dtype)

        def __repr__(self):
            return f"MultiDiscrete({self.nvec.tolist()})"

try:  # pragma: no cover - depends on the installation
    from ray.rllib.env.multi_agent_env import MultiAgentEnv  # type: ignore

    HAVE_RLLIB = True
except Exception:
    HAVE_RLLIB = False

    class MultiAgentEnv:  # the two things the reference uses from RLlib's base class
        def __init__(self, *args, **kwargs):
            pass

        def get_agent_ids(self):
            return set(getattr(self, "agents", []))
