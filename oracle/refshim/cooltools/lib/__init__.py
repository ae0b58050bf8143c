from .. import numutils  # noqa: F401
