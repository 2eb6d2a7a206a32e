"""MiniWorld-ThreeRooms-v0: three connected rooms with assorted objects, no task
(reference envs/threerooms.py)."""
import math

from .._gym import spaces, utils
from ..entity import Ball, Box, ImageFrame, Key, MeshEnt
from ..world import MiniWorldEnv


class ThreeRooms(MiniWorldEnv, utils.EzPickle):
    def __init__(self, **kwargs):
        MiniWorldEnv.__init__(self, max_episode_steps=400, **kwargs)
        utils.EzPickle.__init__(self, **kwargs)
        self.action_space = spaces.Discrete(self.actions.move_forward + 1)

    def _gen_world(self):
        top = self.add_rect_room(min_x=-7, max_x=7, min_z=0.5, max_z=7)
        left = self.add_rect_room(min_x=-7, max_x=-1, min_z=-7, max_z=-0.5)
        right = self.add_rect_room(min_x=1, max_x=7, min_z=-7, max_z=-0.5)
        self.connect_rooms(top, left, min_x=-5.25, max_x=-2.75)
        self.connect_rooms(top, right, min_x=2.75, max_x=5.25)
        self.box = self.place_entity(Box(color="red"))
        self.place_entity(Box(color="green", size=0.6))
        self.entities.append(ImageFrame(pos=[0, 1.35, 7], dir=math.pi / 2, width=1.8, tex_name="logo_mila"))
        self.place_entity(MeshEnt(mesh_name="duckie", height=0.25, static=False))
        self.place_entity(Key(color="blue"))
        self.place_entity(Ball(color="green"))
        self.place_agent()

    device_rule = ("none", 0)

    def device_program(self, prog):
        prog.place(prog.proto(Box(color="red")))
        prog.place(prog.proto(Box(color="green", size=0.6)))
        prog.put(prog.proto(ImageFrame(pos=[0, 1.35, 7], dir=math.pi / 2, width=1.8, tex_name="logo_mila")),
                 pos=[0, 1.35, 7], dir=math.pi / 2, append_only=True)
        prog.place(prog.proto(MeshEnt(mesh_name="duckie", height=0.25, static=False)))
        prog.place(prog.proto(Key(color="blue")))
        prog.place(prog.proto(Ball(color="green")))
        prog.place_agent()

    def step(self, action):
        return super().step(action)
