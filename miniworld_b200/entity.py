"""Scene entities (host side descriptors).

Same public surface as reference `miniworld/entity.py` (Entity :43-121, MeshEnt :124-165,
Box :386-432, Key :435, Ball :445, Agent :455-551) minus every GL call: an entity here is
a plain record (pose, bounding cylinder, appearance) that `pack.py` lowers into the SoA
the CUDA engine consumes.  Appearance is rendered by csrc/raster.cuh, not by `render()`.
"""
import math

import numpy as np

from .assets import ObjMesh, Texture, mesh_ent_dims
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


def _frame_border(sx, hy, hz):
    """The four black border quads shared by ImageFrame and TextFrame (entity.py:223-260, 344-381):
    (normal, four corners) for left, right, top, bottom."""
    return [
        ((0, 0, -1), [(0, +hy, -hz), (+sx, +hy, -hz), (+sx, -hy, -hz), (0, -hy, -hz)]),
        ((0, 0, 1), [(+sx, +hy, +hz), (0, +hy, +hz), (0, -hy, +hz), (+sx, -hy, +hz)]),
        ((0, 1, 0), [(+sx, +hy, +hz), (+sx, +hy, -hz), (0, +hy, -hz), (0, +hy, +hz)]),
        ((0, -1, 0), [(+sx, -hy, -hz), (+sx, -hy, +hz), (0, -hy, +hz), (0, -hy, -hz)]),
    ]


class _QuadFrame(Entity):
    """Wall-mounted frame drawn as a handful of quads.  The reference draws these with immediate
    mode GL (ImageFrame.render / TextFrame.render); here the quads become a small synthetic mesh
    (two triangles per quad, fan order) that the engine's mesh path renders with scale 1."""
    kind = KIND_MESH
    scale = np.float32(1.0)

    @property
    def is_static(self):
        return True

    def _build(self, key, quads):
        """quads: list of (normal, corners[4], uv[4] or None, rgb, texture or None)."""
        V, Nn, UV, C, TT = [], [], [], [], []
        for normal, corners, uvs, rgb, tex in quads:
            for a, b, c in ((0, 1, 2), (0, 2, 3)):
                V.append([corners[a], corners[b], corners[c]])
                Nn.append([normal] * 3)
                UV.append([uvs[a], uvs[b], uvs[c]] if uvs else [(0, 0)] * 3)
                C.append([rgb] * 3)
                TT.append(tex.tex_id if tex is not None else -1)
        self.mesh = ObjMesh.from_arrays(key, V, Nn, UV, C, np.array(TT, np.int32))


class ImageFrame(_QuadFrame):
    """Picture on a wall; `pos` is the middle of the frame, which faces +x before rotation
    (reference entity.py:168-262)."""

    def __init__(self, pos, dir, tex_name, width, depth=0.05):
        super().__init__()
        self.pos, self.dir = pos, dir
        self.tex = Texture.get(tex_name)
        self.width, self.depth = width, depth
        self.height = (float(self.tex.height) / self.tex.width) * self.width
        sx, hz, hy = self.depth, self.width / 2, self.height / 2
        front = ((1, 0, 0), [(sx, +hy, -hz), (sx, +hy, +hz), (sx, -hy, +hz), (sx, -hy, -hz)],
                 [(1, 1), (0, 1), (0, 0), (1, 0)], (1, 1, 1), self.tex)
        border = [(n, c, None, (0, 0, 0), None) for n, c in _frame_border(sx, hy, hz)]
        self._build("imageframe:%s:%r:%r" % (tex_name, width, depth), [front] + border)


class TextFrame(_QuadFrame):
    """Line of character tiles on a wall (reference entity.py:265-383); the glyph textures are
    (re)drawn in `randomize`, i.e. once per reset."""

    def __init__(self, pos, dir, str, height=0.15, depth=0.05):
        super().__init__()
        self.pos, self.dir, self.str = pos, dir, str
        self.depth, self.height = depth, height
        self.width = len(str) * height
        self.mesh = None

    def randomize(self, params, rng):
        self.texs = []
        for ch in self.str:
            try:
                self.texs.append(None if ch == " " else Texture.get("chars/ch_0x%d" % ord(ch), rng))
            except Exception:
                raise ValueError("only alphanumerical characters supported in TextFrame")
        sx, hz, hy = 0.05, self.width / 2, self.height / 2
        quads = []
        for idx, tex in enumerate(self.texs):
            z0 = hz - self.height * (idx + 1)
            z1 = z0 + self.height
            quads.append(((1, 0, 0), [(sx, +hy, z0), (sx, +hy, z1), (sx, -hy, z1), (sx, -hy, z0)],
                          [(1, 1), (0, 1), (0, 0), (1, 0)], (1, 1, 1), tex))
        quads += [(n, c, None, (0, 0, 0), None) for n, c in _frame_border(sx, hy, hz)]
        key = "textframe:%s:%r:%s" % (self.str, self.height, ",".join(str(t.tex_id) if t else "-" for t in self.texs))
        self._build(key, quads)


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
