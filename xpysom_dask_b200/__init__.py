"""B200-native batch-SOM training epoch behind the XPySom API."""
from .xpysom import XPySom  # noqa: F401

__all__ = ["XPySom"]
