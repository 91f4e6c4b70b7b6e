"""Host-side scalar ephemeris services used when spiceypy is not installed."""
from .core import MiniSpice, utc2et, CLIGHT  # noqa: F401
