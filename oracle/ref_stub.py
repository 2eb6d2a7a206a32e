"""TEST INFRASTRUCTURE (oracle) -- never imported by the product path.

Headless import shim for the *unmodified* reference at /root/reference.

The reference (Farama-Foundation/Miniworld) needs pyglet<2 + libGL + gymnasium, none of
which exist in this image.  Every GL entry point it touches is a ctypes call into a
third-party driver; physics / reward / world generation (SURVEY.md section 8a rows R1-R8,
R15, R16) never read anything back from GL.  So we inject fake `pyglet` / `gymnasium`
modules whose GL calls are no-ops and let the reference's own Python execute:
miniworld/miniworld.py, entity.py, math.py, params.py, objmesh.py, opengl.py and envs/*.py
run unmodified from /root/reference (read-only, via sys.path).

This makes the reference itself the oracle of record for pose / dir / reward /
terminated / truncated / RNG order.  Pixels are NOT produced here (render_obs returns
zeros) -- the pixel oracle is oracle/softgl.c.

Only usable where /root/reference exists (the build container); the GPU box uses the
committed fixtures under tests/golden/ produced by oracle/gen_golden.py.
"""
import ctypes
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("MWB_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "miniworld"))


class _Noop:
    """Callable that swallows anything (stands in for a gl* function or GL object)."""

    def __init__(self, name="noop"):
        self._name = name
        self.value = 0
        self.target = 0
        self.id = 0

    def __call__(self, *a, **k):
        return None

    def __getattr__(self, item):
        return _Noop(item)


def _make_gl_module(name):
    mod = types.ModuleType(name)
    ctypes_types = {
        "GLfloat": ctypes.c_float, "GLdouble": ctypes.c_double, "GLubyte": ctypes.c_ubyte,
        "GLuint": ctypes.c_uint, "GLint": ctypes.c_int, "GLushort": ctypes.c_ushort,
        "GLenum": ctypes.c_uint, "GLsizei": ctypes.c_int,
    }
    counter = [0x1000]

    def _getattr(attr):
        if attr.startswith("__"):
            raise AttributeError(attr)
        if attr in ctypes_types:
            return ctypes_types[attr]
        if attr.startswith("GL_"):
            counter[0] += 1
            val = counter[0]
            setattr(mod, attr, val)
            return val
        if attr == "gl_info":
            info = types.SimpleNamespace(have_extension=lambda *_: True)
            return info
        if attr == "glCheckFramebufferStatus":
            fn = lambda *a, **k: getattr(mod, "GL_FRAMEBUFFER_COMPLETE")
        else:
            fn = _Noop(attr)
        setattr(mod, attr, fn)
        return fn

    mod.__getattr__ = _getattr
    return mod


class _FakeImage:
    def __init__(self, path):
        from PIL import Image
        with Image.open(path) as im:
            self.width, self.height = im.size

    def get_texture(self):
        return types.SimpleNamespace(width=self.width, height=self.height, target=0, id=0)

    def get_image_data(self):
        return types.SimpleNamespace(get_data=lambda *_: b"")


def _install_pyglet():
    pyglet = types.ModuleType("pyglet")
    pyglet.options = {}
    gl = _make_gl_module("pyglet.gl")
    pyglet.gl = gl
    image = types.ModuleType("pyglet.image")
    image.load = lambda path, *a, **k: _FakeImage(path)
    image.ImageData = _Noop
    pyglet.image = image
    graphics = types.ModuleType("pyglet.graphics")
    def _vertex_list(count, *attrs):
        """Capture what the reference hands to pyglet (objmesh.py:198-204) so the oracle
        harness can read the mesh arrays back: .attrs['v3f'] etc. are flat float arrays."""
        vl = _Noop("vlist")
        vl.count = count
        vl.attrs = {fmt: data for fmt, data in attrs}
        return vl

    graphics.vertex_list = _vertex_list
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


def _install_gymnasium():
    import numpy as np

    gym = types.ModuleType("gymnasium")

    class Space:
        def __init__(self, shape=None, dtype=None):
            self.shape, self.dtype = shape, dtype

    class Discrete(Space):
        def __init__(self, n):
            super().__init__((), np.int64)
            self.n = int(n)

        def contains(self, x):
            return 0 <= int(x) < self.n

    class Box(Space):
        def __init__(self, low, high, shape=None, dtype=np.float32):
            super().__init__(tuple(shape) if shape is not None else np.shape(low), dtype)
            self.low, self.high = low, high

    class Dict(Space, dict):
        def __init__(self, spaces=None, **kw):
            dict.__init__(self, spaces or {}, **kw)
            Space.__init__(self)

    spaces = types.ModuleType("gymnasium.spaces")
    spaces.Space, spaces.Discrete, spaces.Box, spaces.Dict = Space, Discrete, Box, Dict

    class Env:
        """Seeding semantics of gymnasium.Env.reset (third-party, gymnasium>=0.29.1):
        Generator(PCG64(SeedSequence(seed))) iff a seed is given; otherwise the stream
        continues (lazily created from OS entropy on first use)."""
        metadata = {}
        render_mode = None
        _np_random = None

        def reset(self, *, seed=None, options=None):
            if seed is not None:
                self._np_random = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))

        @property
        def np_random(self):
            if self._np_random is None:
                self._np_random = np.random.Generator(np.random.PCG64(np.random.SeedSequence()))
            return self._np_random

        @np_random.setter
        def np_random(self, v):
            self._np_random = v

        @property
        def unwrapped(self):
            return self

    class EzPickle:
        def __init__(self, *a, **k):
            self._ezpickle_args, self._ezpickle_kwargs = a, k

    utils = types.ModuleType("gymnasium.utils")
    utils.EzPickle = EzPickle
    core = types.ModuleType("gymnasium.core")
    core.ObsType = object
    core.Env = Env
    registry = {}

    def register(id, entry_point=None, **kw):
        registry[id] = (entry_point, kw)

    def make(id, **kw):
        import importlib
        entry, kw0 = registry[id]
        modname, cls = entry.split(":")
        return getattr(importlib.import_module(modname), cls)(**{**kw0.get("kwargs", {}), **kw})

    class Wrapper(Env):
        def __init__(self, env):
            self.env = env

    gym.Env, gym.spaces, gym.utils, gym.core = Env, spaces, utils, core
    gym.register, gym.make, gym.registry = register, make, registry
    gym.Wrapper = gym.ObservationWrapper = gym.ActionWrapper = Wrapper
    gym.logger = types.SimpleNamespace(warn=lambda *a, **k: None)
    for m in (gym, spaces, utils, core):
        sys.modules[m.__name__] = m


_installed = False


def install():
    """Make `import miniworld` resolve to the unmodified reference with GL stubbed out."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    for name in list(sys.modules):
        if name == "miniworld" or name.startswith("miniworld."):
            raise RuntimeError("a module named `miniworld` is already imported")
    if "pyglet" not in sys.modules:
        _install_pyglet()
    if "gymnasium" not in sys.modules:
        _install_gymnasium()
    sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


def make_reference_env(env_id, **kwargs):
    """Instantiate a reference env class by gym id (e.g. 'MiniWorld-FourRooms-v0')."""
    install()
    import contextlib
    import io
    import gymnasium
    import miniworld  # noqa: F401  (registers ids)
    with contextlib.redirect_stdout(io.StringIO()):
        return gymnasium.make(env_id, **kwargs)
