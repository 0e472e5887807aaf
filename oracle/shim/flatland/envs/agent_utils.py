from oracle.trainsim import EnvAgent, Grid4TransitionsEnum  # noqa: F401
