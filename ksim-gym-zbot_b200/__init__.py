"""B200-native batched Z-Bot physics step behind ksim's PhysicsEngine interface."""
from . import model  # noqa: F401

__version__ = "0.1.0"
