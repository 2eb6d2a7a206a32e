"""MiniWorld-WallGap-v0: two outdoor yards joined by a gap in the wall, red box in the far one
(reference envs/wallgap.py)."""
import math

import numpy as np

from .._gym import spaces, utils
from ..entity import Box, MeshEnt
from ..world import MiniWorldEnv


class WallGap(MiniWorldEnv, utils.EzPickle):
    def __init__(self, **kwargs):
        MiniWorldEnv.__init__(self, max_episode_steps=300, **kwargs)
        utils.EzPickle.__init__(self, **kwargs)
        self.action_space = spaces.Discrete(self.actions.move_forward + 1)

    def _gen_world(self):
        yard = dict(wall_tex="brick_wall", floor_tex="asphalt", no_ceiling=True)
        top = self.add_rect_room(min_x=-7, max_x=7, min_z=0.5, max_z=8, **yard)
        bottom = self.add_rect_room(min_x=-7, max_x=7, min_z=-8, max_z=-0.5, **yard)
        self.connect_rooms(top, bottom, min_x=-1.5, max_x=1.5)
        self.box = self.place_entity(Box(color="red"), room=bottom)
        # backdrop
        self.place_entity(MeshEnt(mesh_name="building", height=30), pos=np.array([30, 0, 30]), dir=-math.pi)
        self.place_agent(room=top)

    device_rule = ("goal", 0)

    def device_program(self, prog):
        prog.place(prog.proto(Box(color="red")), room=1)
        prog.put(prog.proto(MeshEnt(mesh_name="building", height=30)), pos=[30, 0, 30], dir=-math.pi)
        prog.place_agent(room=0)

    def step(self, action):
        obs, reward, termination, truncation, info = super().step(action)
        if self.near(self.box):
            reward += self._reward()
            termination = True
        return obs, reward, termination, truncation, info
