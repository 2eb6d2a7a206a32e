import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def _libmwb_built():
    path = os.path.join(ROOT, "miniworld_b200", "libmwb.so")
    nvcc = os.path.exists("/usr/local/cuda/bin/nvcc")
    if nvcc:     # `make` is a no-op when the library is newer than its sources
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "miniworld_b200", "csrc")],
                              stdout=subprocess.DEVNULL)
    return path


@pytest.fixture
def libmwb_path(_libmwb_built):
    """In-tree libmwb.so (built by __graft_entry__.build()): the library every product class loads."""
    from miniworld_b200 import engine
    engine._override_library_for_tests(None)
    return _libmwb_built


@pytest.fixture(scope="session")
def _hostsim_built():
    return subprocess.check_output([os.path.join(ROOT, "tests", "hostsim", "build.sh")]).decode().strip().splitlines()[-1]


@pytest.fixture
def hostsim_path(_hostsim_built):
    """Test-only CPU build of the kernels' inner functions (tests/hostsim/build.sh).  The product classes take no
    library argument: while a test holds this fixture the binding is pointed at the host build through the
    private test seam `engine._override_library_for_tests`."""
    from miniworld_b200 import engine
    engine._override_library_for_tests(_hostsim_built)
    yield _hostsim_built
    engine._override_library_for_tests(None)


@pytest.fixture(scope="session")
def softgl_lib():
    from oracle import softgl
    softgl.build()
    return softgl


_golden_cache = {}


def golden(name):
    """Golden trajectory as a dict of arrays (an NpzFile would decompress an array again on every access)."""
    import numpy as np
    if name not in _golden_cache:
        with np.load(os.path.join(GOLDEN, "traj_%s.npz" % name)) as z:
            _golden_cache[name] = {k: z[k] for k in z.files}
    return _golden_cache[name]
