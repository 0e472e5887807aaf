from oracle.trainsim import TrainState  # noqa: F401
