"""Pixel parity pinned to the reference's own GL stream.

(1) Where /root/reference exists: the UNMODIFIED reference runs under the recording fixed-function GL
    (oracle/gl_record.py); the frames its render_obs() / render_depth() / render_top_view() / get_visible_ents()
    return -- the rasterised stream of the GL calls it made -- must equal, bit for bit, oracle/softgl.py's rendering
    of the package's mirror objects in the same state.  All 23 reference ids (+ the MazeS8 alias), with and without
    domain randomisation.
(2) Everywhere (no reference needed): the kernels' arithmetic compiled for the CPU (tests/hostsim) replays the
    golden trajectories and is compared with the committed frames the reference returned (tests/golden/stream_*.npz).
The same comparison through libmwb.so on a B200 is tests/test_gpu_stream.py.
"""
import numpy as np
import pytest

from helpers import stream_cases, stream_parity
from oracle import ref_stub

needs_reference = pytest.mark.skipif(not ref_stub.reference_available(), reason="needs /root/reference")


def _levels():
    from oracle.stream_check import level_ids
    return level_ids()


@needs_reference
@pytest.mark.parametrize("level", _levels())
def test_reference_gl_stream_equals_mirror(softgl_lib, level):
    from oracle.stream_check import compare
    for dr in (False, True):
        if dr and level == "MiniWorld-Sign-v0":
            continue                                  # Sign fixes domain_rand=False itself (sign.py:88-93)
        n, bad, worst, dbad = compare(level, dr, steps=12)
        assert n == 13 and bad == 0 and dbad == 0, "%s dr=%d: %d / %d frames differ (worst %d LSB), %d depth maps differ" % (
            level, dr, bad, n, worst, dbad)


@needs_reference
@pytest.mark.parametrize("level", ["MiniWorld-Hallway-v0", "MiniWorld-PickupObjects-v0", "MiniWorld-ThreeRooms-v0",
                                   "MiniWorld-Sidewalk-v0", "MiniWorld-Sign-v0"])
def test_reference_other_views_equal_mirror(softgl_lib, level):
    """render_top_view (with the agent marker and its leaked normal), get_visible_ents, a 160 x 120 observation."""
    from oracle.stream_check import Pair
    p = Pair(level, False, obs_width=160, obs_height=120)
    rng = np.random.default_rng(7)
    p.reset(11)
    for t in range(6):
        obs, _, term, trunc, _ = p.step(int(rng.integers(0, p.ref.action_space.n)))
        if term or trunc:
            p.reset(12 + t)
            continue
        obs = obs["obs"] if isinstance(obs, dict) else obs
        assert np.array_equal(obs, p.mirror_frame(160, 120)[0])
        assert np.array_equal(p.ref.render_top_view(), p.mirror_top_view(160, 120))
        assert p.ref_visible() == p.mirror_visible(160, 120)


@needs_reference
def test_reference_human_view_is_16_samples_and_equals_mirror(softgl_lib):
    """render() with render_mode="rgb_array": the reference's vis_fb = FrameBuffer(window_width, window_height, 16)
    (miniworld.py:518); under the recording GL (GL_MAX_SAMPLES = 16) the frame it returns is a 16-sample frame and
    equals the mirror rendered with the D3D 16-sample pattern -- agent view and map view."""
    from oracle.stream_check import Pair
    for view in ("agent", "top"):
        p = Pair("MiniWorld-Hallway-v0", False, render_mode="rgb_array", window_width=160, window_height=120, view=view)
        p.reset(5)
        p.step(2)
        got = p.ref.render()
        fr = ref_stub.recorder.frames[-1]
        assert fr.samples == 16 and got.shape == (120, 160, 3)
        want = p.mirror_frame(160, 120, 16)[0] if view == "agent" else p.mirror_top_view(160, 120, 16)
        assert np.array_equal(got, want)


@needs_reference
def test_reference_own_render_test_holds_on_the_recorded_stream():
    """The one pixel-level statement the reference's test suite makes (tests/test_miniworld.py:17-38, Hallway,
    render_mode="rgb_array"): 0 < mean(obs) < 255, |mean(80x60 observation) - mean(800x600 human view)| < 5, and the
    observation shapes -- evaluated on what the unmodified reference returns under the recording GL."""
    env = ref_stub.make_reference_env("MiniWorld-Hallway-v0", record=True, render_mode="rgb_array")
    for seed in (0, 1):
        env.reset(seed=seed)
        for _ in range(3):
            obs, _, _, _, _ = env.step(0)
        first_obs, _ = env.reset(seed=seed + 10)
        first_render = env.render()
        assert first_render.shape == (600, 800, 3) and ref_stub.recorder.frames[-1].samples == 16
        m0, m1 = first_obs.mean(), first_render.mean()
        assert 0 < m0 < 255
        assert abs(m0 - m1) < 5, (m0, m1)
        second_obs, _, _, _, _ = env.step(0)
        assert first_obs.shape == env.observation_space.shape == second_obs.shape


@needs_reference
def test_reference_light_is_directional():
    """(GLfloat * 4)(*self.light_pos + [1]) with an ndarray light_pos passes THREE values, each + 1, and leaves w = 0
    (miniworld.py:1031, params.py:45-46): LIGHT0 is a directional light along light_pos + 1."""
    env = ref_stub.make_reference_env("MiniWorld-OneRoom-v0", record=True)
    env.reset(seed=3)
    fr = ref_stub.recorder.frames[-1]
    assert isinstance(env.light_pos, np.ndarray)
    assert np.array_equal(fr.light["position"], np.array([1.0, 3.5, 1.0, 0.0], np.float32))


@pytest.mark.parametrize("name", stream_cases())
def test_hostsim_matches_reference_stream_frames(hostsim_path, name):
    # (the CPU build of the kernels is slow on mesh levels: fewer rows there; the GPU test replays every row)
    rows = {"maze_dr": 21, "pickup_160": 3, "pickup": 34, "pickup_dr": 34}.get(name, 89)
    st = stream_parity(name, hostsim_path, max_rows=rows)
    assert st["frames"] >= 4 and st["same"] / st["total"] > 0.995
    assert st["cams"] > 0 and st["cam_exact"] / st["cams"] > 0.98
