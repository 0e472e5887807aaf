from oracle.trainsim import MalfunctionParameters, ParamMalfunctionGen  # noqa: F401
