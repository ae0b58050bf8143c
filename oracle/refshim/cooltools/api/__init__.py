from . import snipping, coverage  # noqa: F401
