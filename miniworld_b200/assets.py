"""Texture / mesh asset store (host side).

Replaces the resource side of reference `opengl.Texture` (opengl.py:103-195: name ->
`name_1.png .. name_9.png` variants, stop at the first gap, domain-rand picks
`rng.integers(0, n)`) and `objmesh.ObjMesh` (objmesh.py:19-216: OBJ/MTL triangle loader
with the recentring rule of :172-186) without any GL: assets are decoded to plain numpy
arrays, registered in a process-wide table whose indices are the texture / mesh ids the
CUDA engine uses.

Sources, in priority order:
  1. a Miniworld resource directory (env MINIWORLD_ASSET_DIR, i.e. the `miniworld/`
     folder of a reference install containing textures/ and meshes/), read with PIL;
  2. the packed store shipped in this repo (assets/pack_v1.npz, built by
     tools/pack_assets.py from those same files) -- used on boxes without the reference.
"""
import math
import os
import threading

import numpy as np

_PACK_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets", "pack_v1.npz")
_lock = threading.Lock()
_pack = None


def _asset_dir():
    d = os.environ.get("MINIWORLD_ASSET_DIR")
    return d if d and os.path.isdir(d) else None


def _load_pack():
    global _pack
    if _pack is None and os.path.exists(_PACK_PATH):
        _pack = np.load(_PACK_PATH)
    return _pack


# ----------------------------------------------------------------------------- textures


class Texture:
    """One decoded texture variant.  `texels` is uint8[H, W, 3], row 0 = TOP of the image
    (the engine flips so that v = 0 is the image bottom, as pyglet's upload does,
    reference opengl.py:158-170)."""

    registry = []          # index == engine texture id
    _by_key = {}
    _variants = {}

    def __init__(self, name, variant, texels):
        self.name = name
        self.variant = variant
        self.texels = np.ascontiguousarray(texels[..., :3], dtype=np.uint8)
        self.height, self.width = self.texels.shape[:2]
        self.tex_id = -1

    @classmethod
    def _read_variant(cls, name, i):
        key = "%s_%d" % (name, i)
        d = _asset_dir()
        if d is not None:
            path = os.path.join(d, "textures", key + ".png")
            if os.path.exists(path):
                from PIL import Image
                with Image.open(path) as im:
                    return np.asarray(im.convert("RGB"))
            return None
        pack = _load_pack()
        if pack is not None and ("tex/" + key) in pack.files:
            return pack["tex/" + key]
        return None

    @classmethod
    def num_variants(cls, name):
        with _lock:
            if name not in cls._variants:
                n = 0
                for i in range(1, 10):
                    if cls._read_variant(name, i) is None:
                        break
                    n += 1
                cls._variants[name] = n
            return cls._variants[name]

    @classmethod
    def family(cls, name):
        """All variants of `name`, registered with consecutive engine ids (the device-side
        domain-rand draw selects `first_id + integers(0, n)`)."""
        n = cls.num_variants(name)
        with _lock:
            if (name, 0) not in cls._by_key:
                for idx in range(n):
                    texels = cls._read_variant(name, idx + 1)
                    tex = Texture(name, idx, texels)
                    tex.tex_id = len(cls.registry)
                    cls.registry.append(tex)
                    cls._by_key[(name, idx)] = tex
            return [cls._by_key[(name, idx)] for idx in range(n)]

    @classmethod
    def from_array(cls, key, texels):
        """Register a one-off texture (a mesh's map_Kd image) under a unique key."""
        with _lock:
            tex = cls._by_key.get((key, 0))
            if tex is None:
                tex = Texture(key, 0, texels)
                tex.tex_id = len(cls.registry)
                cls.registry.append(tex)
                cls._by_key[(key, 0)] = tex
            return tex

    @classmethod
    def get_variant(cls, name, idx):
        fam = cls.family(name)
        if not 0 <= idx < len(fam):
            raise ValueError('failed to load texture "%s" variant %d' % (name, idx + 1))
        return fam[idx]

    @classmethod
    def get(cls, tex_name, rng=None):
        """Reference `Texture.get` semantics (opengl.py:113-145): with an rng the variant
        index is `rng.integers(0, n)` (which draws nothing when n == 1)."""
        n = cls.num_variants(tex_name)
        if n == 0:
            raise ValueError('failed to load textures for name "%s"' % tex_name)
        idx = int(rng.integers(0, n)) if rng is not None else 0
        return cls.get_variant(tex_name, idx)


# ------------------------------------------------------------------------------- meshes


def parse_obj(obj_path, mesh_dir=None):
    """Decode a triangle-only OBJ (+ sibling MTL) into per-face-vertex float32 arrays.

    Follows what the reference loader produces (objmesh.py:36-216): faces stable-sorted
    by material name, colour = material Kd, missing texcoord -> (0, 0), then the
    recentring rule: base to y = 0 and x/z centred with `max` taken as
    verts.max(0).min(0) (sic, objmesh.py:175), extents recomputed afterwards.
    Returns dict(verts, norms, texcs, colors [F,3,*] float32, min_coords, max_coords,
    chunks=[(start, end, texture_path_or_None)]).
    """
    mesh_dir = mesh_dir or os.path.dirname(obj_path)
    stem = os.path.splitext(os.path.basename(obj_path))[0]
    materials = {"": {"Kd": np.array([1, 1, 1])}}
    default_png = os.path.join(mesh_dir, stem + ".png")
    if os.path.exists(default_png):
        materials[""]["map_Kd"] = default_png
    mtl_path = os.path.splitext(obj_path)[0] + ".mtl"
    if os.path.exists(mtl_path):
        cur = None
        with open(mtl_path) as f:
            for raw in f:
                tok = raw.strip().split()
                if not tok or tok[0].startswith("#"):
                    continue
                if tok[0] == "newmtl":
                    cur = materials.setdefault(tok[1], {})
                    cur.clear()
                elif tok[0] == "Kd":
                    cur["Kd"] = np.array([float(v) for v in tok[1:]])
                elif tok[0] == "map_Kd":
                    cur["map_Kd"] = os.path.join(mesh_dir, tok[-1])
    pos, uvs, nrm, faces = [], [], [], []
    mtl = ""
    with open(obj_path) as f:
        for raw in f:
            tok = raw.strip().split()
            if not tok or tok[0].startswith("#"):
                continue
            head, args = tok[0], tok[1:]
            if head == "v":
                pos.append([float(v) for v in args])
            elif head == "vt":
                uvs.append([float(v) for v in args])
            elif head == "vn":
                nrm.append([float(v) for v in args])
            elif head == "usemtl":
                mtl = args[0] if args[0] in materials else ""
            elif head == "f":
                if len(args) != 3:
                    raise ValueError("only triangle faces are supported")
                corners = [[int(i) for i in a.split("/") if i != ""] for a in args]
                faces.append((mtl, corners))
    faces.sort(key=lambda fc: fc[0])
    nf = len(faces)
    verts = np.zeros((nf, 3, 3), np.float32)
    norms = np.zeros((nf, 3, 3), np.float32)
    texcs = np.zeros((nf, 3, 2), np.float32)
    colors = np.zeros((nf, 3, 3), np.float32)
    for fi, (m, corners) in enumerate(faces):
        kd = materials[m].get("Kd", np.array((1, 1, 1))) if materials[m] else np.array((1, 1, 1))
        for ci, idx in enumerate(corners):
            if len(idx) == 3:
                vi, ti, ni = idx
                texcs[fi, ci] = uvs[ti - 1][:2]
            else:
                vi, ni = idx
            verts[fi, ci] = pos[vi - 1]
            norms[fi, ci] = nrm[ni - 1]
            colors[fi, ci] = kd
    lo = verts.min(axis=0).min(axis=0)
    hi_sic = verts.max(axis=0).min(axis=0)
    mid = (lo + hi_sic) / 2
    verts[:, :, 1] -= lo[1]
    verts[:, :, 0] -= mid[0]
    verts[:, :, 2] -= mid[2]
    chunks, start = [], 0
    for fi in range(1, nf + 1):
        if fi == nf or faces[fi][0] != faces[start][0]:
            chunks.append((start, fi, materials[faces[start][0]].get("map_Kd")))
            start = fi
    return dict(verts=verts, norms=norms, texcs=texcs, colors=colors,
                min_coords=verts.min(axis=0).min(axis=0), max_coords=verts.max(axis=0).max(axis=0),
                chunks=chunks)


class ObjMesh:
    """Decoded mesh; `mesh_id` indexes the engine's mesh table."""

    registry = []
    cache = {}

    def __init__(self, name, data):
        self.name = name
        self.verts = np.ascontiguousarray(data["verts"], np.float32)
        self.norms = np.ascontiguousarray(data["norms"], np.float32)
        self.texcs = np.ascontiguousarray(data["texcs"], np.float32)
        self.colors = np.ascontiguousarray(data["colors"], np.float32)
        self.min_coords = np.asarray(data["min_coords"], np.float32)
        self.max_coords = np.asarray(data["max_coords"], np.float32)
        tt = data.get("tri_tex")
        self.tri_tex = np.full(self.verts.shape[0], -1, np.int32) if tt is None else np.ascontiguousarray(tt, np.int32)
        self.mesh_id = -1

    @classmethod
    def from_arrays(cls, key, verts, norms, texcs, colors, tri_tex):
        """Synthetic mesh (ImageFrame / TextFrame quads), cached by `key`."""
        with _lock:
            mesh = cls.cache.get(key)
            if mesh is None:
                verts = np.asarray(verts, np.float32).reshape(-1, 3, 3)
                data = dict(verts=verts, norms=np.asarray(norms, np.float32).reshape(-1, 3, 3),
                            texcs=np.asarray(texcs, np.float32).reshape(-1, 3, 2),
                            colors=np.asarray(colors, np.float32).reshape(-1, 3, 3),
                            min_coords=verts.min(axis=0).min(axis=0), max_coords=verts.max(axis=0).max(axis=0),
                            tri_tex=tri_tex)
                mesh = ObjMesh(key, data)
                mesh.mesh_id = len(cls.registry)
                cls.registry.append(mesh)
                cls.cache[key] = mesh
            return mesh

    @property
    def num_tris(self):
        return self.verts.shape[0]

    @classmethod
    def get(cls, mesh_name):
        with _lock:
            mesh = cls.cache.get(mesh_name)
            if mesh is not None:
                return mesh
            d = _asset_dir()
            data = None
            if d is not None:
                path = os.path.join(d, "meshes", mesh_name + ".obj")
                if os.path.exists(path):
                    data = parse_obj(path)
                    data["tri_tex"] = np.full(len(data["verts"]), -1, np.int32)
                    for start, end, tex_path in data["chunks"]:
                        if tex_path is not None:
                            from PIL import Image
                            with Image.open(tex_path) as im:
                                texels = np.asarray(im.convert("RGB"))
                            data["tri_tex"][start:end] = -2 - len(data.setdefault("_tex", []))
                            data["_tex"].append((os.path.basename(tex_path), texels))
            if data is None:
                pack = _load_pack()
                pre = "mesh/%s/" % mesh_name
                if pack is not None and (pre + "colors") in pack.files:
                    geo = str(pack[pre + "geom"])
                    gpre = "meshgeom/%s/" % geo
                    data = {k: pack[gpre + k] for k in ("verts", "norms", "texcs", "min_coords", "max_coords")}
                    data["colors"] = pack[pre + "colors"]
                    if (gpre + "tri_tex") in pack.files:      # textured chunks: -2 - k refers to meshtex k
                        data["tri_tex"] = pack[gpre + "tri_tex"].copy()
                        data["_tex"] = []
                        k = 0
                        while ("meshtex/%s/%d" % (geo, k)) in pack.files:
                            data["_tex"].append(("%s#%d" % (geo, k), pack["meshtex/%s/%d" % (geo, k)]))
                            k += 1
            if data is None:
                raise ValueError('failed to load mesh "%s"' % mesh_name)
            if data.get("_tex"):                              # resolve chunk textures to engine texture ids
                tt = np.array(data["tri_tex"], np.int32)
                for k, (tex_key, texels) in enumerate(data["_tex"]):
                    _lock.release()
                    try:
                        tex = Texture.from_array("mesh:" + tex_key, texels)
                    finally:
                        _lock.acquire()
                    tt[tt == -2 - k] = tex.tex_id
                data["tri_tex"] = tt
            mesh = ObjMesh(mesh_name, data)
            mesh.mesh_id = len(cls.registry)
            cls.registry.append(mesh)
            cls.cache[mesh_name] = mesh
            return mesh


def mesh_ent_dims(mesh, height):
    """scale / radius of a MeshEnt exactly as reference entity.py:140-147 evaluates them
    under numpy >= 2: max_coords are np.float32 scalars, so `height / sy` and the product
    with the Python-float sqrt are float32 operations."""
    sx, sy, sz = mesh.max_coords
    scale = height / sy
    radius = math.sqrt(sx * sx + sz * sz) * scale
    return scale, radius
