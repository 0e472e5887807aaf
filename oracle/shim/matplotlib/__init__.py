"""oracle shim: cosmetic only (node colours in utils/rail_graph.py:15,50-53)."""


class _CMap:
    def __call__(self, i):
        return (0.0, 0.0, 0.0, 1.0)


class _CMaps(dict):
    def __getitem__(self, k):
        return _CMap()


colormaps = _CMaps()
