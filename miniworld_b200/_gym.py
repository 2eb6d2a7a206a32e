"""gymnasium binding, with a minimal stand-in when gymnasium is not installed.

The reference builds on `gymnasium.Env` / `spaces` / `utils.EzPickle` / `register`
(reference miniworld.py:6-10, envs/__init__.py:44-157).  This image has no gymnasium, so
the few pieces the step path touches are provided here with the same semantics; when the
real package is importable it is used instead.  Seeding follows gymnasium's
`Env.reset(seed=...)`: `Generator(PCG64(SeedSequence(seed)))` iff a seed is passed.
"""
import importlib

import numpy as np

try:  # pragma: no cover - not available in the build image
    import gymnasium as gym
    from gymnasium import spaces, utils
    HAVE_GYMNASIUM = True
except Exception:  # ModuleNotFoundError in this image
    HAVE_GYMNASIUM = False

    class _Space:
        def __init__(self, shape=None, dtype=None):
            self.shape, self.dtype = shape, dtype
            self._rng = None

        def seed(self, seed=None):
            self._rng = np.random.default_rng(seed)

        @property
        def np_random(self):
            if self._rng is None:
                self.seed()
            return self._rng

    class Discrete(_Space):
        def __init__(self, n, start=0):
            super().__init__((), np.int64)
            self.n, self.start = int(n), int(start)

        def sample(self):
            return int(self.start + self.np_random.integers(self.n))

        def contains(self, x):
            return self.start <= int(x) < self.start + self.n

        def __repr__(self):
            return "Discrete(%d)" % self.n

    class Box(_Space):
        def __init__(self, low, high, shape=None, dtype=np.float32):
            shape = tuple(shape) if shape is not None else np.shape(low)
            super().__init__(shape, np.dtype(dtype))
            self.low = np.full(shape, low, dtype=dtype)
            self.high = np.full(shape, high, dtype=dtype)

        def sample(self):
            if np.issubdtype(self.dtype, np.integer):
                return self.np_random.integers(self.low, self.high, endpoint=True, dtype=self.dtype)
            return self.np_random.uniform(self.low, self.high).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and np.all(x >= self.low) and np.all(x <= self.high)

    class Dict(_Space, dict):
        def __init__(self, spaces=None, **kw):
            dict.__init__(self, spaces or {}, **kw)
            _Space.__init__(self)

    class _Spaces:
        Space, Discrete, Box, Dict = _Space, Discrete, Box, Dict

    class _Env:
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
        def np_random(self, value):
            self._np_random = value

        @property
        def unwrapped(self):
            return self

        def close(self):
            pass

    class _EzPickle:
        def __init__(self, *args, **kwargs):
            self._ezpickle_args, self._ezpickle_kwargs = args, kwargs

        def __getstate__(self):
            return {"_ezpickle_args": self._ezpickle_args, "_ezpickle_kwargs": self._ezpickle_kwargs}

        def __setstate__(self, d):
            out = type(self)(*d["_ezpickle_args"], **d["_ezpickle_kwargs"])
            self.__dict__.update(out.__dict__)

    class _Utils:
        EzPickle = _EzPickle

    class _Gym:
        Env = _Env
        registry = {}

        @classmethod
        def register(cls, id, entry_point=None, kwargs=None, **_):
            cls.registry[id] = (entry_point, dict(kwargs or {}))

        @classmethod
        def make(cls, id, **kwargs):
            entry, kw = cls.registry[id]
            if isinstance(entry, str):
                mod, name = entry.split(":")
                entry = getattr(importlib.import_module(mod), name)
            return entry(**{**kw, **kwargs})

    gym, spaces, utils = _Gym, _Spaces, _Utils


def seeded_generator(seed):
    """The generator gymnasium creates for `reset(seed=seed)`."""
    return np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))
