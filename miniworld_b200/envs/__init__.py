"""Level definitions and id registration (reference miniworld/envs/__init__.py:44-157).

All 23 ids of the reference are registered (+ `MiniWorld-MazeS8-v0`, the name BASELINE.json uses for the 8x8
`MiniWorld-Maze-v0` default), and every one of them runs on the batched engine with device-side resets and a
lowered rule: each level class states its `_gen_world()` once more as a `device_program` (CHOICE / UNIFORM /
PLACE / PUT / IFEQ / MAZE ops interpreted per env on the GPU, on that env's numpy-exact stream) and names its
`device_rule` (goal, pickup, sidewalk, sign, health, putnext, none).  The same classes are the drop-in
single-environment API (host world generation, the level's Python `step()`), which is what the reference's
own tests exercise.
"""
from .._gym import gym
from .collecthealth import CollectHealth
from .fourrooms import FourRooms
from .hallway import Hallway
from .maze import Maze, MazeS2, MazeS3, MazeS3Fast
from .oneroom import OneRoom, OneRoomS6, OneRoomS6Fast
from .pickupobjects import PickupObjects
from .putnext import PutNext
from .roomobjects import RoomObjects
from .sidewalk import Sidewalk
from .sign import Sign
from .threerooms import ThreeRooms
from .tmaze import TMaze, TMazeLeft, TMazeRight
from .wallgap import WallGap
from .ymaze import YMaze, YMazeLeft, YMazeRight

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
    # levels outside BASELINE.json's configs
    "MiniWorld-CollectHealth-v0": CollectHealth,
    "MiniWorld-PutNext-v0": PutNext,
    "MiniWorld-RoomObjects-v0": RoomObjects,
    "MiniWorld-Sidewalk-v0": Sidewalk,
    "MiniWorld-Sign-v0": Sign,
    "MiniWorld-ThreeRooms-v0": ThreeRooms,
    "MiniWorld-TMaze-v0": TMaze,
    "MiniWorld-TMazeLeft-v0": TMazeLeft,
    "MiniWorld-TMazeRight-v0": TMazeRight,
    "MiniWorld-WallGap-v0": WallGap,
    "MiniWorld-YMaze-v0": YMaze,
    "MiniWorld-YMazeLeft-v0": YMazeLeft,
    "MiniWorld-YMazeRight-v0": YMazeRight,
}

for _id, _cls in LEVELS.items():
    try:
        gym.register(id=_id, entry_point="%s:%s" % (_cls.__module__, _cls.__name__))
    except Exception:      # already registered (e.g. the reference package is also installed)
        pass

__all__ = sorted({c.__name__ for c in LEVELS.values()})
