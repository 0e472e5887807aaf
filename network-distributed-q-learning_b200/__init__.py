"""switchfl-b200: batched B200 backend for SwitchFL's lockstep hot path (env step + distributed Q-update).

The directory name carries hyphens (it mirrors the reference repository's name), so it is loaded under
the importable alias ``switchfl_b200`` by ``__graft_entry__.load_package()``.
"""
__all__ = ["api", "backend", "cli", "mapgen", "railmap", "sharding"]
