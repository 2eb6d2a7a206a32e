"""MiniWorld-OneRoom-v0 and its S6 / S6Fast variants (reference envs/oneroom.py)."""
from .._gym import spaces, utils
from ..entity import Box
from ..params import DEFAULT_PARAMS
from ..world import MiniWorldEnv
from ._goal import GoalBoxRule


class OneRoom(GoalBoxRule, MiniWorldEnv, utils.EzPickle):
    def __init__(self, size=10, max_episode_steps=180, **kwargs):
        assert size >= 2
        self.size = size
        MiniWorldEnv.__init__(self, max_episode_steps=max_episode_steps, **kwargs)
        utils.EzPickle.__init__(self, size=size, max_episode_steps=max_episode_steps, **kwargs)
        self.action_space = spaces.Discrete(self.actions.move_forward + 1)

    def _layout(self):
        return self.add_rect_room(min_x=0, max_x=self.size, min_z=0, max_z=self.size)

    def _gen_world(self):
        self._layout()
        self.box = self.place_entity(Box(color="red"))
        self.place_agent()

    def device_program(self, prog):
        prog.place(prog.proto(Box(color="red")))
        prog.place_agent()


class OneRoomS6(OneRoom):
    def __init__(self, size=6, max_episode_steps=100, **kwargs):
        super().__init__(size=size, max_episode_steps=max_episode_steps, **kwargs)


fast_params = DEFAULT_PARAMS.no_random()
fast_params.set("forward_step", 0.7)
fast_params.set("turn_step", 45)


class OneRoomS6Fast(OneRoomS6):
    def __init__(self, max_episode_steps=50, params=fast_params, domain_rand=False, **kwargs):
        super().__init__(max_episode_steps=max_episode_steps, params=params, domain_rand=domain_rand, **kwargs)
