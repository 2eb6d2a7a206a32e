"""MiniWorld-Sign-v0: read the sign, go to the object of that colour (reference envs/sign.py;
task from https://arxiv.org/abs/2008.02790).  Observations are dicts {obs, goal}."""
import math

from .._gym import spaces, utils
from ..entity import COLOR_NAMES, Box, Key, MeshEnt, TextFrame
from ..params import DEFAULT_PARAMS
from ..world import MiniWorldEnv


class BigKey(Key):
    """Key scaled up for visibility."""

    def __init__(self, color, size=0.6):
        assert color in COLOR_NAMES
        MeshEnt.__init__(self, mesh_name="key_%s" % color, height=size, static=False)


class Sign(MiniWorldEnv, utils.EzPickle):
    def __init__(self, size=10, max_episode_steps=20, color_index=0, goal=0, **kwargs):
        if color_index not in [0, 1, 2]:
            raise ValueError("Only supported values for color_index are 0, 1, 2.")
        if goal not in [0, 1]:
            raise ValueError("Only supported values for goal are 0, 1.")
        params = DEFAULT_PARAMS.no_random()
        params.set("forward_step", 0.7)
        params.set("turn_step", 45)
        self._size, self._goal, self._color_index = size, goal, color_index
        MiniWorldEnv.__init__(self, params=params, max_episode_steps=max_episode_steps, domain_rand=False, **kwargs)
        utils.EzPickle.__init__(self, size, max_episode_steps, color_index, goal, **kwargs)
        self.observation_space = spaces.Dict(obs=self.observation_space, goal=spaces.Discrete(2))
        self.action_space = spaces.Discrete(self.actions.move_forward + 2)   # + "end episode"

    def set_color_index(self, color_index):
        self._color_index = color_index

    def _gen_world(self):
        s, gap = self._size, 0.25
        top = self.add_rect_room(min_x=0, max_x=s, min_z=0, max_z=s * 0.65)
        left = self.add_rect_room(min_x=0, max_x=s * 3 / 5, min_z=s * 0.65 + gap, max_z=s * 1.3)
        right = self.add_rect_room(min_x=s * 3 / 5, max_x=s, min_z=s * 0.65 + gap, max_z=s * 1.3)
        self.connect_rooms(top, left, min_x=0, max_x=s * 3 / 5)
        self.connect_rooms(left, right, min_z=s * 0.65 + gap, max_z=s * 1.3)
        at = lambda ent, x, z: self.place_entity(ent, pos=(x, 0, z))
        self._objects = [
            (at(Box(color="blue"), 1, 1), at(Box(color="red"), 9, 1), at(Box(color="green"), 9, 5)),
            (at(BigKey(color="blue"), 5, 1), at(BigKey(color="red"), 1, 5), at(BigKey(color="green"), 1, 9)),
        ]
        text = ["BLUE", "RED", "GREEN"][self._color_index]
        self.entities.append(TextFrame(pos=[s, 1.35, s + gap], dir=math.pi, str=text, height=1))
        self.place_agent(min_x=4, max_x=5, min_z=4, max_z=6)

    @property
    def device_rule(self):
        return ("sign", self._color_index | (self._goal << 8))

    @property
    def device_obs_extra(self):
        return {"goal": self._goal}                    # the dict observation's constant entry (sign.py:176)

    def device_program(self, prog):
        s, gap = self._size, 0.25
        spots = [(Box(color="blue"), 1, 1), (Box(color="red"), 9, 1), (Box(color="green"), 9, 5),
                 (BigKey(color="blue"), 5, 1), (BigKey(color="red"), 1, 5), (BigKey(color="green"), 1, 9)]
        for ent, x, z in spots:
            prog.put(prog.proto(ent), pos=(x, 0, z))          # place_entity(pos=...): dir = uniform(-pi, pi)
        text = ["BLUE", "RED", "GREEN"][self._color_index]
        frame = TextFrame(pos=[s, 1.35, s + gap], dir=math.pi, str=text, height=1)
        frame.randomize(self.params, None)                    # builds the glyph quads (no draws: domain_rand is off)
        prog.put(prog.proto(frame), pos=[s, 1.35, s + gap], dir=math.pi, append_only=True)
        prog.place_agent(min_x=4, max_x=5, min_z=4, max_z=6)

    def step(self, action):
        obs, reward, termination, truncation, info = super().step(action)
        if action == self.actions.move_forward + 1:
            termination = True
        for kind, triple in enumerate(self._objects):
            for color_index, obj in enumerate(triple):
                if self.near(obj):
                    termination = True
                    reward = float(color_index == self._color_index and kind == self._goal) * 2 - 1
        return {"obs": obs, "goal": self._goal}, reward, termination, truncation, info

    def reset(self, *, seed=None, options=None):
        obs, info = super().reset(seed=seed, options=options)
        return {"obs": obs, "goal": self._goal}, info
