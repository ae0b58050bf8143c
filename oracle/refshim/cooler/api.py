from coolpuppy_b200.coolio import Cooler as _Cooler


class Cooler(_Cooler):
    """`isinstance(x, cooler.api.Cooler)` is used at coolpup.py:1651."""
