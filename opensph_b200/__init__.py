"""B200-native SPH evaluation engine behind OpenSPH's solver API (hot path only)."""
from . import abi, snapshot  # noqa: F401

__all__ = ["abi", "snapshot"]
