from oracle.trainsim import Waypoint  # noqa: F401
