from . import patches  # noqa: F401
