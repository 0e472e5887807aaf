from oracle.trainsim import get_new_position  # noqa: F401
