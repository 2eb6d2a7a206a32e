"""TEST INFRASTRUCTURE (oracle) -- never imported by the product path.

Headless import shim for the *unmodified* reference at /root/reference.

The reference (Farama-Foundation/Miniworld) needs pyglet<2 + libGL + gymnasium, none of
which exist in this image.  Every GL entry point it touches is a ctypes call into a
third-party driver; physics / reward / world generation (SURVEY.md section 8a rows R1-R8,
R15, R16) never read anything back from GL.  So we inject a fake `gymnasium` and a fake
`pyglet` whose GL is the RECORDING fixed-function GL of oracle/gl_record.py, and let the
reference's own Python execute: miniworld/miniworld.py, entity.py, math.py, params.py,
objmesh.py, opengl.py and envs/*.py run unmodified from /root/reference (read-only, via sys.path).

This makes the reference itself the oracle of record for pose / dir / reward /
terminated / truncated / RNG order, and -- with the recorder active -- for every argument of
every GL call that defines a frame; oracle/softgl.c rasterises that recorded stream.

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
recorder = None      # the recording GL context (oracle/gl_record.RecGL) the reference's GL calls go to


def install():
    """Make `import miniworld` resolve to the unmodified reference with `pyglet.gl` replaced by the recording
    fixed-function GL of oracle/gl_record.py.  `recorder.active` decides whether draw calls are recorded and
    rasterised (the reference's render_* then return the rasterised GL stream) or ignored (physics-only: render_obs
    returns zeros, as with a no-op GL)."""
    global _installed, recorder
    if _installed:
        return
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    for name in list(sys.modules):
        if name == "miniworld" or name.startswith("miniworld."):
            raise RuntimeError("a module named `miniworld` is already imported")
    if "pyglet" not in sys.modules:
        from oracle import gl_record
        recorder = gl_record.install()
    if "gymnasium" not in sys.modules:
        _install_gymnasium()
    sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


def make_reference_env(env_id, record=False, **kwargs):
    """Instantiate a reference env class by gym id (e.g. 'MiniWorld-FourRooms-v0').  `record` switches the
    process-wide GL recorder on (frames are rasterised from the reference's GL stream) or off (physics only)."""
    install()
    if recorder is not None:
        recorder.active = bool(record)
    import contextlib
    import io
    import gymnasium
    import miniworld  # noqa: F401  (registers ids)
    with contextlib.redirect_stdout(io.StringIO()):
        return gymnasium.make(env_id, **kwargs)
