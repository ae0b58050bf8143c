"""Stand-in for ``cooler`` (unpinned in the reference's requirements.txt).

Backed by :mod:`coolpuppy_b200.coolio`, whose matrix/extent/offset/bins
semantics restate cooler 0.9 (`Cooler.matrix(sparse=True, balance=name).fetch`
mirrors the lower triangle and multiplies ``w[row] * w[col] * count``).
"""
from . import api
from .api import Cooler

__version__ = "0.0-refshim"
