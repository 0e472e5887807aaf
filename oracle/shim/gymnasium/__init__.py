"""oracle shim: the gymnasium surface the reference touches."""
from gymnasium.spaces import Space  # noqa: F401
from gymnasium import spaces  # noqa: F401
