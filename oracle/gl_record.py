"""TEST INFRASTRUCTURE (oracle) -- never imported by the product path.

A *recording* fixed-function OpenGL for the unmodified reference (Farama-Foundation/Miniworld @
c660156): a fake `pyglet.gl` that keeps the GL 1.x state the reference programs -- matrix stacks,
current colour / normal / texcoord, enables, bound textures and their uploaded texels, LIGHT0,
display lists, framebuffer objects, occlusion queries -- and turns every primitive the reference
submits into triangles carrying exactly the arguments it passed:

    MiniWorldEnv.render_obs / render_top_view     miniworld.py:1088-1221  glClearColor, gluPerspective /
                                                  glOrtho, gluLookAt / glLoadMatrixf arguments
    _render_static / _render_world                miniworld.py:1019-1086  glLightfv, display list 1, draw order
    Room._render                                  miniworld.py:401-434    glColor / glNormal / glTexCoord / glVertex
    Box / MeshEnt / ImageFrame / TextFrame / Agent .render   entity.py:150-161, 196-262, 304-383, 409-432, 518-539
    drawBox                                       opengl.py:460-503
    ObjMesh.render                                objmesh.py:280-292      vertex-list arrays, per-chunk texture
    Texture.load                                  opengl.py:147-184       the RGBA bytes handed to glTexImage2D
    FrameBuffer                                   opengl.py:197-435       sample count, resolve blit, glReadPixels
    get_visible_ents                              miniworld.py:1238-1333  GL_ANY_SAMPLES_PASSED queries

When the reference reads a framebuffer back (glReadPixels) or asks for a query result, the recorded
stream of that frame is rasterised by oracle/softgl.c and the result is written where the reference
asked for it -- so the UNMODIFIED reference's render_obs() / render_depth() / render_top_view() /
get_visible_ents() / render() return images whose every input (geometry, attributes, transforms, light,
camera, clear colour, texels, draw order, GL state leaks) is what the reference submitted.  What stays
a restatement is only what a GL *driver* does with that stream (DESIGN.md "pixel spec"): float32
transform arithmetic, sample positions, quad split, LOD, mip filter, resolve rounding.

Stream -> triangle conventions (the same ones oracle/softgl.py applies to the package's mirror objects, so the
two can be compared bit for bit):
  * every glVertex3f / glNormal3f / glTexCoord2f / glColor3f / glTranslatef / glRotatef / glScalef argument is
    rounded to float32 (they are GLfloat parameters; ctypes rounds the Python float);
  * the model part of the model-view stack (everything after gluLookAt / glLoadMatrixf) is applied to the
    object-space vertex in float32, one rounding per operation, innermost transform first:
    glRotatef(a, 0, 1, 0): x' = x c + z s, z' = z c - x s with c, s = float32(cos / sin(float64(a) * pi / 180)),
    a being the float32 angle in degrees the call received; glScalef: v * s; glTranslatef: v + t.
    Normals go through the inverse transpose (R n, then * float32(1 / s)) and are not renormalised;
  * GL_POLYGON and each GL_QUADS quad are split as the fan (0,1,2), (0,2,3), ...;
  * after an array draw the current colour / normal are those of the last array element.
"""
import ctypes
import math
import sys
import types

import numpy as np

f32 = np.float32

# real enum values where arithmetic on them matters; everything else gets a unique number on demand
_ENUMS = {
    "GL_COLOR_BUFFER_BIT": 0x4000, "GL_DEPTH_BUFFER_BIT": 0x0100, "GL_LINES": 1, "GL_LINE_STRIP": 3, "GL_TRIANGLES": 4,
    "GL_QUADS": 7, "GL_POLYGON": 9, "GL_TEXTURE_2D": 0x0DE1, "GL_LIGHTING": 0x0B50, "GL_LIGHT0": 0x4000,
    "GL_PROJECTION": 0x1701, "GL_MODELVIEW": 0x1700, "GL_COMPILE": 0x1300, "GL_FRAMEBUFFER": 0x8D40,
    "GL_READ_FRAMEBUFFER": 0x8CA8, "GL_DRAW_FRAMEBUFFER": 0x8CA9, "GL_POSITION": 0x1203, "GL_AMBIENT": 0x1200,
    "GL_DIFFUSE": 0x1201, "GL_RGB": 0x1907, "GL_RGBA": 0x1908, "GL_DEPTH_COMPONENT": 0x1902,
    "GL_UNSIGNED_BYTE": 0x1401, "GL_UNSIGNED_SHORT": 0x1403, "GL_FLOAT": 0x1406, "GL_TEXTURE_2D_MULTISAMPLE": 0x9100,
    "GL_FRAMEBUFFER_COMPLETE": 0x8CD5, "GL_COLOR_ATTACHMENT0": 0x8CE0, "GL_DEPTH_ATTACHMENT": 0x8D00,
    "GL_ANY_SAMPLES_PASSED": 0x8C2F, "GL_QUERY_RESULT": 0x8866, "GL_CULL_FACE": 0x0B44, "GL_DEPTH_TEST": 0x0B71,
    "GL_MULTISAMPLE": 0x809D, "GL_COLOR_MATERIAL": 0x0B57, "GL_RENDERBUFFER": 0x8D41, "GL_LINEAR": 0x2601,
    "GL_NEAREST": 0x2600, "GL_SMOOTH": 0x1D01, "GL_FRONT_AND_BACK": 0x0408, "GL_AMBIENT_AND_DIFFUSE": 0x1602,
}
_CTYPES = {
    "GLfloat": ctypes.c_float, "GLdouble": ctypes.c_double, "GLubyte": ctypes.c_ubyte, "GLuint": ctypes.c_uint,
    "GLint": ctypes.c_int, "GLushort": ctypes.c_ushort, "GLenum": ctypes.c_uint, "GLsizei": ctypes.c_int,
}
E = types.SimpleNamespace(**{k[3:]: v for k, v in _ENUMS.items()})


def _val(x):
    """ctypes scalar / byref / plain number -> int"""
    if hasattr(x, "_obj"):
        x = x._obj
    return int(x.value) if hasattr(x, "value") else int(x)


def _addr(p):
    """address behind a ctypes pointer / array / byref / int"""
    if p is None:
        return 0
    if isinstance(p, int):
        return p
    if hasattr(p, "_obj"):
        return ctypes.addressof(p._obj)
    if isinstance(p, ctypes.Array):
        return ctypes.addressof(p)
    return ctypes.cast(p, ctypes.c_void_p).value or 0


class Frame:
    """Everything drawn into one framebuffer between a glClear and the read-back."""

    def __init__(self, width, height, samples, clear):
        self.width, self.height, self.samples = width, height, samples
        self.clear = clear                     # float32 rgba of glClearColor
        self.proj = None                       # ('perspective', fovy, aspect, near, far) | ('ortho', l, r, b, t, n, f)
        self.view = None                       # ('lookat', 9 doubles) | ('loadmatrix', 16 float32)
        self.light = None                      # dict(position, ambient, diffuse) float32[4] each, world space
        self.batches = []                      # dicts of object-space triangle arrays + model ops + texture + query
        self.stale_light = False               # drawn under a light issued for another view: colours not modelled
        self.result = None                     # (rgb u8[H,W,3] top row first, codes u16[H,W]) once rasterised
        self.query_hits = None

    def num_tris(self):
        return sum(len(b["tex"]) for b in self.batches)


class RecGL:
    """The GL context: state + recording.  One instance stands for the process-wide context the reference creates
    with its hidden pyglet window."""

    def __init__(self, max_samples=16):
        self.max_samples = max_samples         # what glGetIntegerv(GL_MAX_SAMPLES) reports (opengl.py:223-231)
        self.active = True                     # False: draw calls / read-backs are ignored (physics-only runs: fast)
        self.enabled = set()
        self.matrix_mode = E.MODELVIEW
        self.stacks = {E.MODELVIEW: [[]], E.PROJECTION: [[]]}     # stack of op lists
        self.color = np.array([1, 1, 1, 1], f32)
        self.normal = np.array([0, 0, 1], f32)
        self.texcoord = np.array([0, 0], f32)
        self.clear_color = np.array([0, 0, 0, 0], f32)
        self.light = {"position": np.array([0, 0, 1, 0], f32), "ambient": np.array([0, 0, 0, 1], f32),
                      "diffuse": np.array([1, 1, 1, 1], f32), "stack": ()}
        self.textures = {}                     # id -> dict(width, height, rgba (bytes, rows bottom-up) | None, path, samples)
        self.bound_tex = {}                    # target -> id
        self.fbos = {0: {"color": None, "frame": None, "resolved": None}}
        self.draw_fbo = self.read_fbo = 0
        self.renderbuffers = {}
        self.lists = {}
        self.compiling = None
        self.prim = None                       # (mode, [vertex records]) between glBegin and glEnd
        self.query = None
        self.queries = {}                      # id -> (frame, index)
        self.next_id = 1
        self.frames = []                       # every frame ever cleared, newest last (bounded)
        self.unknown_calls = set()

    # ------------------------------------------------------------------ helpers
    def _new_id(self):
        self.next_id += 1
        return self.next_id

    def _top(self, mode=None):
        return self.stacks[self.matrix_mode if mode is None else mode][-1]

    def _frame(self):
        fb = self.fbos[self.draw_fbo]
        if fb.get("frame") is None:
            raise RuntimeError("draw call before glClear on framebuffer %d" % self.draw_fbo)
        return fb["frame"]

    def _fbo_dims(self, fbo):
        tex = self.textures.get(self.fbos[fbo].get("color"))
        if tex is None:
            raise RuntimeError("framebuffer %d has no colour attachment" % fbo)
        return tex["width"], tex["height"], max(1, tex.get("samples", 1))

    # ------------------------------------------------------------------ display lists
    def record_or_run(self, name, fn, args):
        if not self.active and name in _DRAW_ONLY:
            return None
        if self.compiling is not None and name not in _IMMEDIATE:
            self.compiling.append((fn, args))
            return None
        return fn(*args)

    def glNewList(self, lst, mode):
        assert mode == E.COMPILE
        self.compiling = []
        self._compiling_id = _val(lst)

    def glEndList(self):
        self.lists[self._compiling_id] = self.compiling
        self.compiling = None

    def glCallList(self, lst):
        if not self.active:
            return
        for fn, args in self.lists.get(_val(lst), []):
            fn(*args)

    def glDeleteLists(self, lst, n):
        for k in range(_val(lst), _val(lst) + _val(n)):
            self.lists.pop(k, None)

    # ------------------------------------------------------------------ state
    def glEnable(self, cap):
        self.enabled.add(_val(cap))

    def glDisable(self, cap):
        self.enabled.discard(_val(cap))

    def glColor3f(self, r, g, b):
        self.color = np.array([r, g, b, 1.0], f32)

    def glNormal3f(self, x, y, z):
        self.normal = np.array([x, y, z], f32)

    def glTexCoord2f(self, u, v):
        self.texcoord = np.array([u, v], f32)

    def glClearColor(self, r, g, b, a):
        self.clear_color = np.array([r, g, b, a], f32)

    def glLightfv(self, light, pname, params):
        assert _val(light) == E.LIGHT0
        v = np.array(list(params)[:4], f32)
        key = {E.POSITION: "position", E.AMBIENT: "ambient", E.DIFFUSE: "diffuse"}[_val(pname)]
        self.light = dict(self.light)
        self.light[key] = v
        if key == "position":
            # GL stores the position in eye space = model-view * params; we keep the stack it was issued under and
            # require (at rasterisation time) that it is the bare view transform, so params are world coordinates
            self.light["stack"] = tuple(self._top(E.MODELVIEW))

    # ------------------------------------------------------------------ matrices
    def glMatrixMode(self, mode):
        self.matrix_mode = _val(mode)

    def glLoadIdentity(self):
        self.stacks[self.matrix_mode][-1] = []

    def glPushMatrix(self):
        st = self.stacks[self.matrix_mode]
        st.append(list(st[-1]))

    def glPopMatrix(self):
        self.stacks[self.matrix_mode].pop()

    def gluPerspective(self, fovy, aspect, near, far):
        self._top().append(("perspective", float(fovy), float(aspect), float(near), float(far)))

    def glOrtho(self, l, r, b, t, n, f):
        self._top().append(("ortho", float(l), float(r), float(b), float(t), float(n), float(f)))

    def gluLookAt(self, *a):
        self._top().append(("lookat",) + tuple(float(v) for v in a))

    def glLoadMatrixf(self, m):
        self.stacks[self.matrix_mode][-1] = [("loadmatrix",) + tuple(float(f32(v)) for v in list(m)[:16])]

    def glTranslatef(self, x, y, z):
        self._top().append(("translate", f32(x), f32(y), f32(z)))

    def glScalef(self, x, y, z):
        self._top().append(("scale", f32(x), f32(y), f32(z)))

    def glRotatef(self, a, x, y, z):
        self._top().append(("rotate", f32(a), f32(x), f32(y), f32(z)))

    # ------------------------------------------------------------------ textures
    def glGenTextures(self, n, ref):
        arr = ref._obj if hasattr(ref, "_obj") else ref
        tid = self._new_id()
        self.textures[tid] = {"width": 0, "height": 0, "rgba": None, "path": None, "samples": 1}
        arr.value = tid

    def glBindTexture(self, target, tex):
        self.bound_tex[_val(target)] = _val(tex)

    def glTexImage2D(self, target, level, ifmt, w, h, border, fmt, typ, data):
        tid = self.bound_tex.get(_val(target), 0)
        t = self.textures.setdefault(tid, {"path": None, "samples": 1})
        t["width"], t["height"] = _val(w), _val(h)
        if data is not None and _val(typ) == E.UNSIGNED_BYTE and _val(fmt) == E.RGBA:
            raw = bytes(data) if not isinstance(data, (bytes, bytearray)) else bytes(data)
            assert len(raw) >= t["width"] * t["height"] * 4
            t["rgba"] = raw[:t["width"] * t["height"] * 4]
            t["internal"] = _val(ifmt)
        else:
            t["rgba"] = None

    def glTexImage2DMultisample(self, target, samples, ifmt, w, h, fixed):
        tid = self.bound_tex.get(_val(target), 0)
        t = self.textures.setdefault(tid, {"path": None})
        t.update(width=_val(w), height=_val(h), samples=_val(samples), rgba=None)

    # ------------------------------------------------------------------ framebuffers
    def glGetIntegerv(self, pname, out):
        if _val(pname) == 0x8D57:              # GL_MAX_SAMPLES
            (out._obj if hasattr(out, "_obj") else out).value = self.max_samples

    def glGenFramebuffers(self, n, ref):
        fid = self._new_id()
        self.fbos[fid] = {"color": None, "frame": None, "resolved": None}
        ref._obj.value = fid

    def glGenRenderbuffers(self, n, ref):
        ref._obj.value = self._new_id()

    def glBindFramebuffer(self, target, fbo):
        t, f = _val(target), _val(fbo)
        if t in (E.FRAMEBUFFER, E.DRAW_FRAMEBUFFER):
            self.draw_fbo = f
        if t in (E.FRAMEBUFFER, E.READ_FRAMEBUFFER):
            self.read_fbo = f

    def glFramebufferTexture2D(self, target, attachment, textarget, tex, level):
        if _val(attachment) == E.COLOR_ATTACHMENT0:
            self.fbos[self.draw_fbo]["color"] = _val(tex)

    def glCheckFramebufferStatus(self, target):
        return E.FRAMEBUFFER_COMPLETE

    def glClear(self, mask):
        if not (_val(mask) & E.COLOR_BUFFER_BIT) or not self.active:
            return
        if self.draw_fbo == 0:
            return                             # the window's own buffer (human mode): not a framebuffer we model
        w, h, s = self._fbo_dims(self.draw_fbo)
        fr = Frame(w, h, s, self.clear_color.copy())
        self.fbos[self.draw_fbo]["frame"] = fr
        self.frames.append(fr)
        del self.frames[:-8]

    def glBlitFramebuffer(self, sx0, sy0, sx1, sy1, dx0, dy0, dx1, dy1, mask, filt):
        # FrameBuffer.resolve (opengl.py:345-374): multisample -> single-sample, colour then depth
        src = self.fbos[self.read_fbo]["frame"]
        self.fbos[self.draw_fbo]["resolved"] = src

    def glReadPixels(self, x, y, w, h, fmt, typ, ptr):
        if not self.active:
            return
        fr = self.fbos[self.read_fbo].get("resolved")
        if fr is None:
            fr = self.fbos[self.read_fbo].get("frame")
        if fr is None:
            raise RuntimeError("glReadPixels from a framebuffer nothing was rendered into")
        rgb, codes = self.rasterise(fr)
        w, h = _val(w), _val(h)
        assert (_val(x), _val(y), w, h) == (0, 0, fr.width, fr.height)
        if _val(fmt) == E.RGB and _val(typ) == E.UNSIGNED_BYTE:
            if fr.stale_light:
                raise NotImplementedError("colour read-back of a frame lit from a stale light position")
            out = np.ascontiguousarray(rgb[::-1])          # glReadPixels starts at the lower-left corner
        elif _val(fmt) == E.DEPTH_COMPONENT and _val(typ) == E.UNSIGNED_SHORT:
            out = np.ascontiguousarray(codes[::-1])
        else:
            raise NotImplementedError("glReadPixels format")
        ctypes.memmove(_addr(ptr), out.ctypes.data, out.nbytes)

    # ------------------------------------------------------------------ queries
    def glGenQueries(self, n, ids):
        for k in range(_val(n)):
            ids[k] = self._new_id()

    def glBeginQuery(self, target, qid):
        if not self.active:
            return
        fr = self._frame()
        self.query = _val(qid)
        self.queries[self.query] = (fr, None)

    def glEndQuery(self, target):
        self.query = None

    def glGetQueryObjectuiv(self, qid, pname, out):
        if not self.active:
            return
        fr, _ = self.queries[_val(qid)]
        self.rasterise(fr)
        out[0] = 1 if fr.query_hits.get(_val(qid), False) else 0

    def glDeleteQueries(self, n, ids):
        pass

    # ------------------------------------------------------------------ primitives
    def _state_key(self):
        tex = -1
        if E.TEXTURE_2D in self.enabled:
            tex = self.bound_tex.get(E.TEXTURE_2D, 0)
            if tex == 0:
                tex = -1
        return tex

    def _submit(self, pos, nrm, uv, rgb, tex):
        """pos [T,3,3] object space float32, etc.  Snapshot the transform / light / projection state."""
        if not self.active:
            return
        fr = self._frame()
        mv = self._top(E.MODELVIEW)
        proj = tuple(self._top(E.PROJECTION))
        if not mv or mv[0][0] not in ("lookat", "loadmatrix"):
            raise NotImplementedError("model-view stack does not start with gluLookAt / glLoadMatrixf: %r" % (mv[:1],))
        view, model = mv[0], tuple(mv[1:])
        if len(proj) != 1:
            raise NotImplementedError("projection stack %r" % (proj,))
        if fr.proj is None:
            fr.proj, fr.view = proj[0], view
        elif fr.proj != proj[0] or fr.view != view:
            raise NotImplementedError("camera changed inside a frame")
        if E.LIGHTING not in self.enabled or E.LIGHT0 not in self.enabled or E.COLOR_MATERIAL not in self.enabled:
            raise NotImplementedError("primitive drawn without LIGHTING + LIGHT0 + COLOR_MATERIAL")
        light = {k: self.light[k].copy() for k in ("position", "ambient", "diffuse")}
        if tuple(self.light["stack"]) != (view,):
            # GL keeps the light's EYE-space position from the model-view it was issued under; drawing under another
            # view (get_visible_ents after a top-view render: it never calls the display list) lights the scene from a
            # stale direction.  Coverage / depth / queries are unaffected; colours of such a frame are not modelled.
            fr.stale_light = True
        if fr.light is None:
            fr.light = light
        elif any(not np.array_equal(fr.light[k], light[k]) for k in light):
            raise NotImplementedError("light changed inside a frame")
        fr.batches.append({"pos": pos, "nrm": nrm, "uv": uv, "rgb": rgb, "tex": np.full(len(pos), tex, np.int32),
                           "model": model, "query": self.query})

    def glBegin(self, mode):
        self.prim = (_val(mode), [])

    def glVertex3f(self, x, y, z):
        if not self.active:
            return
        self.prim[1].append((np.array([x, y, z], f32), self.normal.copy(), self.texcoord.copy(), self.color[:3].copy()))

    def glEnd(self):
        mode, vs = self.prim
        self.prim = None
        if not self.active or mode in (E.LINES, E.LINE_STRIP):
            return                             # debug lines (Entity.draw_bound, drawAxes): never reached by the API paths
        n = len(vs)
        if mode == E.POLYGON:
            idx = [(0, k, k + 1) for k in range(1, n - 1)]
        elif mode == E.QUADS:
            idx = [(q + a, q + b, q + c) for q in range(0, n - 3, 4) for a, b, c in ((0, 1, 2), (0, 2, 3))]
        elif mode == E.TRIANGLES:
            idx = [(k, k + 1, k + 2) for k in range(0, n - 2, 3)]
        else:
            raise NotImplementedError("primitive mode %d" % mode)
        if not idx:
            return
        ii = np.asarray(idx)
        arr = lambda j: np.stack([v[j] for v in vs]).astype(f32)[ii]
        self._submit(arr(0), arr(1), arr(2), arr(3), self._state_key())

    def draw_vertex_list(self, vl, mode):
        """pyglet VertexList.draw(GL_TRIANGLES) with v3f / t2f / n3f / c3f arrays (objmesh.py:198-204, 290)."""
        assert _val(mode) == E.TRIANGLES
        get = lambda key, k: np.asarray(vl.attrs[key], f32).reshape(-1, 3, k)
        pos, uv, nrm, rgb = get("v3f", 3), get("t2f", 2), get("n3f", 3), get("c3f", 3)
        self._submit(pos, nrm, uv, rgb, self._state_key())
        self.normal = nrm[-1, -1].copy()       # de-facto behaviour: the last array element stays current
        self.color = np.array(list(rgb[-1, -1]) + [1.0], f32)

    # ------------------------------------------------------------------ stream -> softgl
    def texture_table(self):
        """(list of uint8[H,W,3] top-row-first images, {gl texture id: index}) of every texture that was uploaded
        with texels -- i.e. exactly the bytes the reference handed to glTexImage2D (opengl.py:161-171)."""
        imgs, index = [], {}
        for tid in sorted(self.textures):
            t = self.textures[tid]
            if t.get("rgba") is None:
                continue
            a = np.frombuffer(t["rgba"], np.uint8).reshape(t["height"], t["width"], 4)
            index[tid] = len(imgs)
            imgs.append(np.ascontiguousarray(a[::-1, :, :3]))      # GL row 0 = bottom; GL_RGB internal format drops alpha
        return imgs, index

    @staticmethod
    def world_triangles(batch):
        """Apply the recorded model transforms in float32 (module docstring)."""
        pos, nrm = batch["pos"].astype(f32), batch["nrm"].astype(f32)
        for op in reversed(batch["model"]):
            if op[0] == "rotate":
                a, ax, ay, az = op[1:]
                if (float(ax), float(ay), float(az)) != (0.0, 1.0, 0.0):
                    raise NotImplementedError("rotation axis %r" % (op,))
                rad = float(a) * math.pi / 180
                c, s = f32(math.cos(rad)), f32(math.sin(rad))
                x, z = pos[..., 0].copy(), pos[..., 2].copy()
                pos[..., 0] = x * c + z * s
                pos[..., 2] = z * c - x * s
                nx, nz = nrm[..., 0].copy(), nrm[..., 2].copy()
                nrm[..., 0] = nx * c + nz * s
                nrm[..., 2] = nz * c - nx * s
            elif op[0] == "scale":
                sx, sy, sz = op[1:]
                if not (sx == sy == sz):
                    raise NotImplementedError("non-uniform scale")
                pos = pos * sx
                nrm = nrm * f32(f32(1.0) / sx)
            elif op[0] == "translate":
                pos = pos + np.array(op[1:], f32)
            else:
                raise NotImplementedError(op[0])
        return pos.astype(f32), nrm.astype(f32)

    def frame_arrays(self, fr):
        """The frame's triangles in world space + texture indices into texture_table()."""
        imgs, index = self.texture_table()
        P, Nn, UV, RGB, TX, Q = [], [], [], [], [], []
        for b in fr.batches:
            p, n = self.world_triangles(b)
            P.append(p); Nn.append(n); UV.append(b["uv"]); RGB.append(b["rgb"])
            TX.append(np.array([index[t] if t >= 0 else -1 for t in b["tex"]], np.int32))
            Q.append(np.full(len(p), -1 if b["query"] is None else b["query"], np.int64))
        cat = lambda xs, shape: (np.concatenate(xs) if xs else np.zeros(shape)).astype(xs[0].dtype if xs else f32)
        return (cat(P, (0, 3, 3)), cat(Nn, (0, 3, 3)), cat(UV, (0, 3, 2)), cat(RGB, (0, 3, 3)),
                cat(TX, (0,)).astype(np.int32), cat(Q, (0,)).astype(np.int64), imgs)

    def rasterise(self, fr):
        if fr.result is not None:
            return fr.result
        from oracle import softgl
        pos, nrm, uv, rgb, tx, q, imgs = self.frame_arrays(fr)
        ts = self._texset(imgs)
        qids = sorted(set(int(v) for v in q if v >= 0))
        qmap = {v: k for k, v in enumerate(qids)}
        query = np.array([qmap.get(int(v), -1) for v in q], np.int32) if qids else None
        flags = np.zeros(max(1, len(qids)), np.uint8)
        cam = {"clear": fr.clear, "light": fr.light, "proj": fr.proj, "view": fr.view}
        out, _, codes = softgl.run_stream(cam, ts, (pos, nrm, uv, rgb, tx), fr.width, fr.height, fr.samples,
                                          query=query, query_out=flags if qids else None)
        fr.query_hits = {v: bool(flags[k]) for v, k in qmap.items()}
        fr.result = (out, codes)
        return fr.result

    def _texset(self, imgs):
        from oracle import softgl
        key = len(imgs)
        if getattr(self, "_ts_key", None) != key:
            if getattr(self, "_ts", None) is not None:
                self._ts.close()
            self._ts, self._ts_key = softgl.TextureSet(imgs), key
        return self._ts


_DRAW_ONLY = {"glBegin", "glEnd", "glVertex3f", "glNormal3f", "glTexCoord2f", "glColor3f", "glPushMatrix", "glPopMatrix",
              "glTranslatef", "glRotatef", "glScalef", "vlist.draw"}     # (state that outlives a frame -- binds, enables, light -- is always tracked)
_IMMEDIATE = {"glNewList", "glEndList", "glGenTextures", "glGenFramebuffers", "glGenRenderbuffers", "glGenQueries",
              "glGetIntegerv", "glCheckFramebufferStatus", "glReadPixels", "glGetQueryObjectuiv", "glDeleteLists",
              "glDeleteQueries", "glTexImage2D", "glTexImage2DMultisample", "glFramebufferTexture2D"}


class _RecTexture:
    def __init__(self, ctx, path, width, height):
        self.target, self.id = E.TEXTURE_2D, ctx._new_id()
        self.width, self.height = width, height
        ctx.textures[self.id] = {"width": width, "height": height, "rgba": None, "path": path, "samples": 1}


class _RecImage:
    """pyglet.image.load(path): size + the RGBA bytes pyglet hands out for a positive pitch (rows bottom-up)."""

    def __init__(self, ctx, path):
        from PIL import Image
        self.ctx, self.path = ctx, path
        with Image.open(path) as im:
            self._rgba = np.asarray(im.convert("RGBA"), np.uint8)
        self.height, self.width = self._rgba.shape[:2]

    def get_texture(self):
        return _RecTexture(self.ctx, self.path, self.width, self.height)

    def get_image_data(self):
        img = self

        class _Data:
            def get_data(self, fmt, pitch):
                assert fmt == "RGBA" and pitch == img.width * 4
                return np.ascontiguousarray(img._rgba[::-1]).tobytes()

        return _Data()


class _RecVertexList:
    def __init__(self, ctx, count, attrs):
        self.ctx, self.count = ctx, count
        self.attrs = {fmt: np.array(data, f32) for fmt, data in attrs}

    def draw(self, mode):
        self.ctx.record_or_run("vlist.draw", self.ctx.draw_vertex_list, (self, mode))

    def delete(self):
        pass


class _Noop:
    def __init__(self, name="noop"):
        self._name, self.value, self.target, self.id = name, 0, 0, 0

    def __call__(self, *a, **k):
        return None

    def __getattr__(self, item):
        return _Noop(item)


def install(max_samples=16):
    """Inject the recording `pyglet` into sys.modules (before the reference is imported).  Returns the context."""
    ctx = RecGL(max_samples)
    pyglet = types.ModuleType("pyglet")
    pyglet.options = {}
    gl = types.ModuleType("pyglet.gl")
    counter = [0x100000]

    def _getattr(attr):
        if attr.startswith("__"):
            raise AttributeError(attr)
        if attr in _CTYPES:
            return _CTYPES[attr]
        if attr.startswith("GL_"):
            if attr in _ENUMS:
                val = _ENUMS[attr]
            else:
                counter[0] += 1
                val = counter[0]
            setattr(gl, attr, val)
            return val
        if attr == "gl_info":
            return types.SimpleNamespace(have_extension=lambda *_: True)
        if attr == "Config":
            return _Noop("Config")
        impl = getattr(RecGL, attr, None)
        if impl is not None:
            bound = getattr(ctx, attr)
            fn = (lambda *a, _b=bound, _n=attr: ctx.record_or_run(_n, _b, a))
        else:
            ctx.unknown_calls.add(attr)
            fn = _Noop(attr)
        setattr(gl, attr, fn)
        return fn

    gl.__getattr__ = _getattr
    pyglet.gl = gl
    image = types.ModuleType("pyglet.image")
    image.load = lambda path, *a, **k: _RecImage(ctx, path)
    image.ImageData = _Noop
    pyglet.image = image
    graphics = types.ModuleType("pyglet.graphics")
    graphics.vertex_list = lambda count, *attrs: _RecVertexList(ctx, count, attrs)
    pyglet.graphics = graphics
    window = types.ModuleType("pyglet.window")
    window.Window = lambda *a, **k: _Noop("window")
    window.key = _Noop("key")
    pyglet.window = window
    text = types.ModuleType("pyglet.text")
    text.Label = lambda *a, **k: _Noop("label")
    pyglet.text = text
    pyglet.app = _Noop("app")
    pyglet.clock = _Noop("clock")
    for m in (pyglet, gl, image, graphics, window, text):
        sys.modules[m.__name__] = m
    pyglet._mwb_recorder = ctx
    return ctx
