"""The C-ABI library loads and exports exactly what include/mwb.h declares (no GPU needed)."""
import ctypes
import os
import re

import pytest

from miniworld_b200 import engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "mwb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mwb_[a-z_]+)\s*\(", src)))


def test_header_functions_all_exported(libmwb_path):
    lib = ctypes.CDLL(libmwb_path)
    names = declared_functions()
    assert len(names) >= 17
    for name in names:
        assert hasattr(lib, name), "libmwb.so does not export %s" % name
    assert sorted(engine.EXPORTS) == names, "engine.py binds a different set than mwb.h declares"


def test_struct_mirrors_match(libmwb_path):
    engine.load_library()                 # raises EngineError on any sizeof mismatch


def test_no_cpu_fallback(libmwb_path):
    """Without a CUDA device the product refuses to run instead of falling back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(engine.EngineError, match="no CUDA device"):
        engine.Engine(1)


def test_missing_library_is_loud(tmp_path):
    engine._override_library_for_tests(str(tmp_path / "libmwb.so"))
    try:
        with pytest.raises(engine.EngineError, match="not found"):
            engine.load_library()
    finally:
        engine._override_library_for_tests(None)


def test_product_classes_take_no_library_argument():
    """The only library the package loads is its own libmwb.so: no constructor accepts a path (the host build of the
    kernels used by CPU tests is reachable only through the private test seam)."""
    import inspect
    from miniworld_b200.batched import BatchedMiniWorld
    from miniworld_b200.dist import ShardedMiniWorld
    from miniworld_b200.world import MiniWorldEnv
    for cls in (BatchedMiniWorld, MiniWorldEnv, engine.Engine, engine.SingleEnvEngine, engine.SharedDeviceBuffer, ShardedMiniWorld):
        params = inspect.signature(cls.__init__).parameters
        assert not any("lib" in name for name in params), (cls.__name__, list(params))
    assert list(inspect.signature(engine.load_library).parameters) == []
