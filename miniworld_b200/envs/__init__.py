"""Level definitions and id registration (reference miniworld/envs/__init__.py:44-157).

Only the levels on BASELINE.json's configured path live here; `MiniWorld-MazeS8-v0` is the
8x8 maze (the reference's `MiniWorld-Maze-v0` default) under the name BASELINE.json uses.
"""
from .._gym import gym
from .fourrooms import FourRooms
from .hallway import Hallway
from .maze import Maze, MazeS2, MazeS3, MazeS3Fast
from .oneroom import OneRoom, OneRoomS6, OneRoomS6Fast
from .pickupobjects import PickupObjects

LEVELS = {
    "MiniWorld-Hallway-v0": Hallway,
    "MiniWorld-OneRoom-v0": OneRoom,
    "MiniWorld-OneRoomS6-v0": OneRoomS6,
    "MiniWorld-OneRoomS6Fast-v0": OneRoomS6Fast,
    "MiniWorld-FourRooms-v0": FourRooms,
    "MiniWorld-Maze-v0": Maze,
    "MiniWorld-MazeS8-v0": Maze,
    "MiniWorld-MazeS2-v0": MazeS2,
    "MiniWorld-MazeS3-v0": MazeS3,
    "MiniWorld-MazeS3Fast-v0": MazeS3Fast,
    "MiniWorld-PickupObjects-v0": PickupObjects,
}

for _id, _cls in LEVELS.items():
    try:
        gym.register(id=_id, entry_point="%s:%s" % (_cls.__module__, _cls.__name__))
    except Exception:      # already registered (e.g. the reference package is also installed)
        pass

__all__ = sorted({c.__name__ for c in LEVELS.values()})
