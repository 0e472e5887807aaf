from oracle.trainsim import RailEnv, RailEnvActions  # noqa: F401
