"""TEST INFRASTRUCTURE ONLY."""
from . import Interval, Position  # noqa: F401


class NoLinearProjectionException(Exception):
    pass
