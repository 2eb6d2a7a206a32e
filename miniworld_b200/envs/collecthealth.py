"""MiniWorld-CollectHealth-v0: health drains every step, picking up a med-kit restores it
(reference envs/collecthealth.py)."""
from .._gym import utils
from ..entity import MeshEnt
from ..world import MiniWorldEnv


class CollectHealth(MiniWorldEnv, utils.EzPickle):
    def __init__(self, size=16, **kwargs):
        assert size >= 2
        self.size = size
        MiniWorldEnv.__init__(self, max_episode_steps=1000, **kwargs)
        utils.EzPickle.__init__(self, size, **kwargs)

    def _gen_world(self):
        self.add_rect_room(min_x=0, max_x=self.size, min_z=0, max_z=self.size, wall_tex="cinder_blocks", floor_tex="slime")
        for _ in range(18):
            self.box = self.place_entity(MeshEnt(mesh_name="medkit", height=0.40, static=False))
        self.place_agent()
        self.health = 100

    device_rule = ("health", 0)
    device_info = {"health": ("counter",)}             # info["health"] = self.health (kept by the rule on the device)

    def device_program(self, prog):
        kit = prog.proto(MeshEnt(mesh_name="medkit", height=0.40, static=False))
        for _ in range(18):
            prog.place(kit)
        prog.place_agent()

    def step(self, action):
        obs, reward, termination, truncation, info = super().step(action)
        self.health -= 2
        if action == self.actions.pickup and self.agent.carrying:
            kit = self.agent.carrying                 # respawn the kit somewhere else
            self.entities.remove(kit)
            self.place_entity(kit)
            self.agent.carrying = None
            self.health = 100
        if self.health > 0:
            reward = 2
        else:
            reward = -100
            termination = True
        info["health"] = self.health
        return obs, reward, termination, truncation, info
