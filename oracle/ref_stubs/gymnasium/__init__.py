"""Stub of the tiny `gymnasium` surface the reference env touches (test infra only)."""
from . import spaces  # noqa: F401


class Env:
    def __init__(self, *args, **kwargs):
        pass
