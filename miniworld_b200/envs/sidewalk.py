"""MiniWorld-Sidewalk-v0: walk along the sidewalk to the red box; stepping into the street ends
the episode (reference envs/sidewalk.py)."""
import math

import numpy as np

from .._gym import spaces, utils
from ..entity import Box, MeshEnt
from ..world import MiniWorldEnv


class Sidewalk(MiniWorldEnv, utils.EzPickle):
    def __init__(self, **kwargs):
        MiniWorldEnv.__init__(self, max_episode_steps=150, **kwargs)
        utils.EzPickle.__init__(self, **kwargs)
        self.action_space = spaces.Discrete(self.actions.move_forward + 1)

    def _gen_world(self):
        walk = self.add_rect_room(min_x=-3, max_x=0, min_z=0, max_z=12, wall_tex="brick_wall",
                                  floor_tex="concrete_tiles", no_ceiling=True)
        self.street = self.add_rect_room(min_x=0, max_x=6, min_z=-80, max_z=80, floor_tex="asphalt", no_ceiling=True)
        self.connect_rooms(walk, self.street, min_z=0, max_z=12)
        self.place_entity(MeshEnt(mesh_name="building", height=30), pos=np.array([30, 0, 30]), dir=-math.pi)
        for i in range(1, walk.max_z // 2):           # traffic cones along the kerb
            self.place_entity(MeshEnt(mesh_name="cone", height=0.75), pos=np.array([1, 0, 2 * i]))
        self.box = self.place_entity(Box(color="red"), room=walk, min_z=walk.max_z - 2, max_z=walk.max_z)
        self.place_agent(room=walk, min_z=0, max_z=1.5)

    @property
    def device_rule(self):
        n_cones = len(range(1, int(self.rooms[0].max_z // 2)))
        return ("sidewalk", (1 + n_cones) | (1 << 8))     # box slot after the building and the cones; street = room 1

    def device_program(self, prog):
        walk = self.rooms[0]
        prog.put(prog.proto(MeshEnt(mesh_name="building", height=30)), pos=[30, 0, 30], dir=-math.pi)
        cone = prog.proto(MeshEnt(mesh_name="cone", height=0.75))
        for i in range(1, int(walk.max_z // 2)):
            prog.put(cone, pos=[1, 0, 2 * i])                  # dir drawn: uniform(-pi, pi)
        prog.place(prog.proto(Box(color="red")), room=0, min_z=walk.max_z - 2, max_z=walk.max_z)
        prog.place_agent(room=0, min_z=0, max_z=1.5)

    def step(self, action):
        obs, reward, termination, truncation, info = super().step(action)
        if self.street.point_inside(self.agent.pos):
            reward = 0
            termination = True
        if self.near(self.box):
            reward += self._reward()
            termination = True
        return obs, reward, termination, truncation, info
