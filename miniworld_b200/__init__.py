"""miniworld_b200: B200-native batched MiniWorld step engine."""
from . import envs  # noqa: F401  (registers the MiniWorld-* ids)
from .params import DEFAULT_PARAMS, DomainParams  # noqa: F401
from .world import MiniWorldEnv, Room  # noqa: F401

__version__ = "0.1.0"
