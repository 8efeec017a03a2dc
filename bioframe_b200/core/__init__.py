"""Host-side mirror of `bioframe.core` for the pieces the hot path needs."""
from . import arrops, checks, specs  # noqa: F401
