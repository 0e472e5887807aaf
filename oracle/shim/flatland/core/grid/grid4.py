from oracle.trainsim import Grid4TransitionsEnum  # noqa: F401
