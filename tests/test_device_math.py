"""Device math helpers, compiled for the host (tests/hostsim), against libm and numpy."""
import ctypes as C
import math

import numpy as np
import pytest


@pytest.fixture(scope="module")
def hs(_hostsim_built):
    return C.CDLL(_hostsim_built)


@pytest.mark.parametrize("scale", [0.12, 0.8, 2.4, 10.0, 1000.0, 1e5])
def test_sincos_matches_glibc_bit_for_bit(hs, scale):
    """csrc/libm_sincos.cuh reproduces this image's libm (what math.cos/math.sin call in the
    reference's dir_vec / right_vec, entity.py:95-113)."""
    rng = np.random.default_rng(int(scale * 1000))
    x = rng.uniform(-scale, scale, 400_000)
    s, c = np.empty_like(x), np.empty_like(x)
    hs.hs_sincos(x.ctypes.data_as(C.c_void_p), len(x), s.ctypes.data_as(C.c_void_p), c.ctypes.data_as(C.c_void_p))
    want_s = np.array([math.sin(v) for v in x])
    want_c = np.array([math.cos(v) for v in x])
    assert np.array_equal(s.view(np.uint64), want_s.view(np.uint64))
    assert np.array_equal(c.view(np.uint64), want_c.view(np.uint64))


def test_sincos_on_headings_the_engine_produces(hs):
    """dir = dir0 + k * 15 deg accumulated in float64, the values the physics kernel sees."""
    rng = np.random.default_rng(7)
    d = rng.uniform(-math.pi, math.pi, 2000)
    xs = []
    for _ in range(200):
        d = d + rng.choice([-1.0, 1.0], d.shape) * (15 * (math.pi / 180))
        xs.append(d.copy())
    x = np.concatenate(xs)
    s, c = np.empty_like(x), np.empty_like(x)
    hs.hs_sincos(x.ctypes.data_as(C.c_void_p), len(x), s.ctypes.data_as(C.c_void_p), c.ctypes.data_as(C.c_void_p))
    assert np.array_equal(s, np.array([math.sin(v) for v in x]))
    assert np.array_equal(c, np.array([math.cos(v) for v in x]))


@pytest.mark.parametrize("seed", [0, 1, 42, 1000, 2 ** 40 + 7])
def test_pcg64_stream_matches_numpy(hs, seed):
    """csrc/np_rng.cuh draws what numpy's Generator(PCG64(SeedSequence(seed))) draws, in any
    interleaving of random / uniform / integers (incl. the buffered 32-bit path, n == 1)."""
    from miniworld_b200.engine import rng_state_of
    gen = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))
    rec = rng_state_of(seed)
    plan = np.random.default_rng(seed + 1)
    n = 3000
    kind = plan.integers(0, 3, n).astype(np.int32)
    a, b = np.zeros(n), np.zeros(n)
    want = np.zeros(n)
    for i in range(n):
        if kind[i] == 0:
            want[i] = gen.random()
        elif kind[i] == 1:
            lo, hi = sorted(plan.uniform(-40, 40, 2))
            a[i], b[i] = lo, hi - lo
            want[i] = gen.uniform(lo, hi)
        else:
            m = int(plan.choice([1, 2, 3, 4, 6, 8, 11, 127, 1000, 2 ** 20 + 3]))
            a[i] = m
            want[i] = gen.integers(0, m)
    st = np.array([rec["state_hi"], rec["state_lo"], rec["inc_hi"], rec["inc_lo"]], np.uint64)
    has32, cache = C.c_int(int(rec["has_uint32"])), C.c_uint32(int(rec["uinteger"]))
    out = np.zeros(n)
    hs.hs_rng_draws(st.ctypes.data_as(C.c_void_p), C.byref(has32), C.byref(cache), kind.ctypes.data_as(C.c_void_p),
                    a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), n, out.ctypes.data_as(C.c_void_p))
    assert np.array_equal(out, want)
    after = gen.bit_generator.state
    assert (int(st[0]) << 64 | int(st[1])) == after["state"]["state"]
    assert has32.value == after["has_uint32"] and cache.value == after["uinteger"]
