"""Stub of RLlib's MultiAgentEnv base class (test infra only)."""


class MultiAgentEnv:
    def __init__(self, *args, **kwargs):
        pass

    def get_agent_ids(self):
        return set(getattr(self, "agents", []))
