from oracle.trainsim import RailGridTransitionMap  # noqa: F401
