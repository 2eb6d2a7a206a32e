"""MiniWorld-FourRooms-v0: four 6x6 rooms joined by 2 m openings (reference envs/fourrooms.py)."""
from .._gym import spaces, utils
from ..entity import Box
from ..world import MiniWorldEnv
from ._goal import GoalBoxRule


class FourRooms(GoalBoxRule, MiniWorldEnv, utils.EzPickle):
    def __init__(self, **kwargs):
        MiniWorldEnv.__init__(self, max_episode_steps=250, **kwargs)
        utils.EzPickle.__init__(self, **kwargs)
        self.action_space = spaces.Discrete(self.actions.move_forward + 1)

    def _layout(self):
        nw = self.add_rect_room(min_x=-7, max_x=-1, min_z=1, max_z=7)
        ne = self.add_rect_room(min_x=1, max_x=7, min_z=1, max_z=7)
        se = self.add_rect_room(min_x=1, max_x=7, min_z=-7, max_z=-1)
        sw = self.add_rect_room(min_x=-7, max_x=-1, min_z=-7, max_z=-1)
        self.connect_rooms(nw, ne, min_z=3, max_z=5, max_y=2.2)
        self.connect_rooms(ne, se, min_x=3, max_x=5, max_y=2.2)
        self.connect_rooms(se, sw, min_z=-5, max_z=-3, max_y=2.2)
        self.connect_rooms(sw, nw, min_x=-5, max_x=-3, max_y=2.2)

    def _gen_world(self):
        self._layout()
        self.box = self.place_entity(Box(color="red"))
        self.place_agent()

    def device_program(self, prog):
        prog.place(prog.proto(Box(color="red")))
        prog.place_agent()
