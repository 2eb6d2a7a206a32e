"""Scene entities (host side descriptors).

Same public surface as reference `miniworld/entity.py` (Entity :43-121, MeshEnt :124-165,
Box :386-432, Key :435, Ball :445, Agent :455-551) minus every GL call: an entity here is
a plain record (pose, bounding cylinder, appearance) that `pack.py` lowers into the SoA
the CUDA engine consumes.  Appearance is rendered by csrc/raster.cuh, not by `render()`.
"""
import math

import numpy as np

from .assets import ObjMesh, mesh_ent_dims
from .math import X_VEC, Y_VEC, Z_VEC, gen_rot_matrix

COLORS = {
    "red": np.array([1.0, 0.0, 0.0]),
    "green": np.array([0.0, 1.0, 0.0]),
    "blue": np.array([0.0, 0.0, 1.0]),
    "purple": np.array([0.44, 0.15, 0.76]),
    "yellow": np.array([1.00, 1.00, 0.00]),
    "grey": np.array([0.39, 0.39, 0.39]),
}
COLOR_NAMES = sorted(COLORS)

# engine entity kinds (csrc/mwb_types.h)
KIND_NONE, KIND_BOX, KIND_MESH, KIND_AGENT = 0, 1, 2, 3


class Entity:
    kind = KIND_NONE

    def __init__(self):
        self.pos = None      # world position, floor level for most entities
        self.dir = None      # heading, radians
        self.radius = 0      # bounding cylinder
        self.height = 0

    def randomize(self, params, rng):
        """Domain-randomisation hook, called once per reset in entity-list order."""

    def step(self, delta_time):
        pass

    @property
    def dir_vec(self):
        return np.array([math.cos(self.dir), 0, -math.sin(self.dir)])

    @property
    def right_vec(self):
        return np.array([math.sin(self.dir), 0, math.cos(self.dir)])

    @property
    def is_static(self):
        return False


class MeshEnt(Entity):
    """Entity drawn from an OBJ mesh scaled to `height`."""
    kind = KIND_MESH

    def __init__(self, mesh_name, height, static=True):
        super().__init__()
        self.static = static
        self.mesh_name = mesh_name
        self.mesh = ObjMesh.get(mesh_name)
        self.scale, self.radius = mesh_ent_dims(self.mesh, height)
        self.height = height

    @property
    def is_static(self):
        return self.static


class Box(Entity):
    kind = KIND_BOX

    def __init__(self, color, size=0.8):
        super().__init__()
        if type(size) is int or type(size) is float:
            size = np.array([size, size, size])
        size = np.array(size)
        sx, sy, sz = size
        self.color = color
        self.size = size
        self.radius = math.sqrt(sx * sx + sz * sz) / 2
        self.height = sy
        self.color_vec = COLORS[color]

    def randomize(self, params, rng):
        self.color_vec = np.clip(COLORS[self.color] + params.sample(rng, "obj_color_bias"), 0, 1)


class Key(MeshEnt):
    def __init__(self, color):
        assert color in COLOR_NAMES
        super().__init__(mesh_name="key_%s" % color, height=0.35, static=False)


class Ball(MeshEnt):
    def __init__(self, color, size=0.6):
        assert color in COLOR_NAMES
        super().__init__(mesh_name="ball_%s" % color, height=size, static=False)


class Agent(Entity):
    kind = KIND_AGENT

    def __init__(self):
        super().__init__()
        self.cam_height = 1.5      # metres above the floor
        self.cam_pitch = 0         # degrees, positive looks up
        self.cam_fov_y = 60        # degrees
        self.cam_fwd_disp = 0
        self.radius = 0.4
        self.height = 1.6
        self.carrying = None

    @property
    def cam_pos(self):
        disp = np.dot(np.array([self.cam_fwd_disp, self.cam_height, 0]), gen_rot_matrix(Y_VEC, self.dir))
        return self.pos + disp

    @property
    def cam_dir(self):
        d = np.dot(X_VEC, gen_rot_matrix(Z_VEC, self.cam_pitch * math.pi / 180))
        return np.dot(d, gen_rot_matrix(Y_VEC, self.dir))

    def randomize(self, params, rng):
        params.sample_many(rng, self, ["cam_height", "cam_fwd_disp", "cam_pitch", "cam_fov_y"])
