"""Stand-in for the ``cooltools`` (>=0.5.2) functions on the reference path."""
from . import numutils  # noqa: F401

__version__ = "0.0-refshim"
