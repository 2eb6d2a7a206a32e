"""Kernel logic on the CPU: the MWB_DEV functions of csrc/ compiled by g++ (tests/hostsim)
replay reference trajectories and render frames.  This checks the kernels' arithmetic where
no GPU exists; the `-m gpu` tests run the real kernels through libmwb.so."""
import numpy as np
import pytest

from conftest import golden
from helpers import CASES, make_env, run_trajectory


@pytest.mark.parametrize("name,steps,n", [("hallway", 60, 16), ("oneroom", 60, 16), ("fourrooms", 80, 16),
                                          ("fourrooms_dr", 60, 8), ("pickup", 120, 8), ("pickup_dr", 60, 8),
                                          ("mazes3", 40, 4), ("maze_dr", 12, 3)])
def test_physics_and_reset_bit_exact(hostsim_path, name, steps, n):
    run_trajectory(name, golden(name), hostsim_path, steps=steps, n=n, check_every=10)


def test_render_matches_oracle(hostsim_path, softgl_lib):
    from miniworld_b200.assets import Texture
    from miniworld_b200.envs import LEVELS
    g = golden("fourrooms")
    env = make_env("fourrooms", g, hostsim_path, n=3, want_depth=True)
    N = env.num_envs
    obs = np.zeros((N, 60, 80, 3), np.uint8)
    depth = np.zeros((N, 60, 80, 1), np.float32)
    env.engine.render(obs=obs, depth=depth)
    ts = softgl_lib.TextureSet([t.texels for t in Texture.registry])
    for i in range(N):
        ref = LEVELS["MiniWorld-FourRooms-v0"](device=None)
        ref.reset(seed=1000 + i)
        rgb, d = softgl_lib.render(ref, ts, lambda tex: tex.tex_id)
        assert np.abs(rgb.astype(int) - obs[i].astype(int)).max() <= 1
        assert np.array_equal(d, depth[i])
        assert 0 < rgb.mean() < 255
    ts.close()
    env.close()


@pytest.mark.parametrize("name,n", [("mazes3", 6), ("maze_dr", 4)])
def test_device_maze_geometry_equals_host_world(hostsim_path, name, n):
    """csrc/maze.cuh (recursive backtracker on the env's numpy stream + translated templates)
    builds the same rooms / quads / collision segments as the Python `_gen_world()`."""
    from miniworld_b200 import pack
    from miniworld_b200.envs import LEVELS
    level, dr = CASES[name]
    g = golden(name)
    env = make_env(name, g, hostsim_path, n=n)
    assert env.device_reset and env.maze_template is not None
    for i in range(n):
        host = LEVELS[level](device=None, domain_rand=dr)
        host.reset(seed=1000 + i)
        want = pack.pack_geometry(host)
        got = env.engine.get_geometry(i)
        for w, d, what in zip(want, got, ("rooms", "quads", "segs")):
            assert len(w) == len(d), (what, len(w), len(d))
            for field in w.dtype.names:
                if field != "reserved":
                    assert np.array_equal(w[field], d[field]), (what, field, i)
    env.close()


def test_device_maze_long_rollout_with_resets(hostsim_path):
    run_trajectory("mazes3", golden("mazes3"), hostsim_path, steps=300, n=8, check_every=25)


def test_maze_truncation_at_max_episode_steps(hostsim_path):
    """1600 steps of the 8 x 8 maze: step_count reaches max_episode_steps = 1536 (maze.py:49), the step is truncated
    (reward 0, terminated False) and the next one resets -- bit for bit as the reference did."""
    g = golden("maze_long")
    assert g["truncated"].sum() >= 1 and g["step_count"].max() == 1536
    run_trajectory("maze_long", g, hostsim_path, n=1, check_every=64)


@pytest.mark.parametrize("name", ["tmaze", "ymaze_dr", "roomobjs", "putnext_dr", "pickup", "wallgap", "sidewalk_dr",
                                  "collecthealth", "collecthealth_pick", "threerooms_dr", "sign"])
def test_single_env_levels_follow_reference(hostsim_path, name):
    """Levels outside the batched configs (and PickupObjects for the carry path) through the
    N = 1 engine with the level's own Python rule."""
    from helpers import run_single_env_trajectory
    run_single_env_trajectory(name, golden(name), hostsim_path, envs=2, steps=40 if name == "pickup" else 100)


def test_human_view_16_samples_matches_oracle(hostsim_path, softgl_lib):
    """render() = the reference's vis_fb frame, 16 samples per pixel (miniworld.py:518): kernels' arithmetic on the CPU
    with the 16-sample pattern vs the oracle, agent view and map view."""
    from miniworld_b200.assets import Texture
    from miniworld_b200.envs import Hallway
    for view in ("agent", "top"):
        env = Hallway(render_mode="rgb_array", window_width=120, window_height=90, view=view)
        env.reset(seed=3)
        env.step(2)
        frame = env.render()
        ts = softgl_lib.TextureSet([t.texels for t in Texture.registry])
        want = (softgl_lib.render(env, ts, lambda tex: tex.tex_id, 120, 90, 16)[0] if view == "agent"
                else softgl_lib.render_top_view(env, ts, lambda tex: tex.tex_id, 120, 90, 16))
        ts.close()
        d = np.abs(frame.astype(int) - want.astype(int))
        assert frame.shape == (90, 120, 3) and d.max() <= 1 and (d == 0).mean() > 0.995
        env.close()


def test_render_mode_and_wrappers_on_host_sim(hostsim_path):
    """reference tests/test_miniworld.py:17-64 (render vs obs mean, wrapper shapes), kernels on the CPU."""
    from miniworld_b200.envs import Hallway
    from miniworld_b200.wrappers import GreyscaleWrapper, PyTorchObsWrapper, StochasticActionWrapper
    env = Hallway(render_mode="rgb_array", window_width=200, window_height=150)
    env.reset(seed=0)
    for _ in range(3):
        obs, _, _, _, _ = env.step(2)
        frame = env.render()
        assert frame.shape == (150, 200, 3) and abs(obs.mean() - frame.mean()) < 5
    env.close()
    w = PyTorchObsWrapper(Hallway())
    assert w.reset()[0].shape == (3, 80, 60) == tuple(w.observation_space.shape)
    g = GreyscaleWrapper(Hallway())
    assert g.reset()[0].shape == (60, 80, 1)
    s = StochasticActionWrapper(Hallway(), prob=0.5)
    s.reset(seed=1)
    s.step(0)
    for e in (w, g, s):
        e.close()


@pytest.mark.parametrize("level", ["MiniWorld-Hallway-v0", "MiniWorld-FourRooms-v0", "MiniWorld-PickupObjects-v0",
                                   "MiniWorld-ThreeRooms-v0"])
def test_top_view_and_visible_ents_match_oracle(hostsim_path, softgl_lib, level):
    """render_top_view / get_visible_ents (reference miniworld.py:1088-1175, 1238-1333): kernels' arithmetic on
    the CPU vs the immediate-mode oracle (depth-buffered draws with GL_ANY_SAMPLES_PASSED bookkeeping)."""
    from miniworld_b200.assets import Texture
    from miniworld_b200.envs import LEVELS
    env = LEVELS[level]()
    seen = 0
    for seed in (3, 4):
        env.reset(seed=seed)
        for _ in range(4):
            env.step(env.action_space.sample())
        ts = softgl_lib.TextureSet([t.texels for t in Texture.registry])
        top = env.render_top_view()
        ref = softgl_lib.render_top_view(env, ts, lambda tex: tex.tex_id)
        assert np.abs(top.astype(int) - ref.astype(int)).max() <= 1
        assert (top == ref).mean() > 0.999 and 0 < top.mean() < 255
        vis = env.get_visible_ents()
        assert vis == softgl_lib.visible_ents(env, ts, lambda tex: tex.tex_id)
        seen += len(vis)
        ts.close()
    img, scale = env.render_top_view(return_scale=True)
    assert img.shape == (60, 80, 3) and scale["x_scale"] > 0 and scale["z_scale"] > 0
    env.close()


def test_device_action_noise_follows_wrapper(hostsim_path):
    from helpers import noise_parity
    noise_parity(hostsim_path)
    noise_parity(hostsim_path, n=2, steps=40, prob=0.3, random_action=1)


@pytest.mark.parametrize("level", ["MiniWorld-FourRooms-v0", "MiniWorld-MazeS3-v0", "MiniWorld-PickupObjects-v0"])
def test_snapshot_restore_resumes_bit_exact(hostsim_path, level):
    from helpers import snapshot_roundtrip
    snapshot_roundtrip(level, hostsim_path, n=4, before=20, after=30)


def test_fused_observation_layouts(hostsim_path):
    from helpers import obs_format_parity
    obs_format_parity(hostsim_path, n=2, steps=2)


@pytest.mark.parametrize("name", ["tmaze", "ymaze_dr", "wallgap", "sidewalk_dr", "threerooms_dr", "roomobjs", "sign", "collecthealth", "collecthealth_pick", "putnext_dr"])
def test_lowered_extra_levels_bit_exact(hostsim_path, name):
    """TMaze / YMaze (branching placement, polygon rooms), WallGap / ThreeRooms / Sidewalk (fixed-pose entities,
    meshes, the street rule) through the batched engine with device-side resets vs the reference trajectories."""
    g = golden(name)
    env = make_env(name, g, hostsim_path, n=2)
    assert env.device_reset
    env.close()
    run_trajectory(name, g, hostsim_path, steps=300 if name == "collecthealth_pick" else 150, check_every=10)


@pytest.mark.parametrize("level,dr", [("MiniWorld-PutNext-v0", True), ("MiniWorld-Sign-v0", False),
                                      ("MiniWorld-TMaze-v0", False)])
def test_batched_frames_equal_single_env(hostsim_path, level, dr):
    from helpers import batched_equals_single_env
    batched_equals_single_env(level, hostsim_path, n=2, steps=2, domain_rand=dr)


def test_every_level_views_match_oracle(hostsim_path, softgl_lib):
    """Every registered level (domain randomisation on where the level allows it): first-person frame, depth map,
    top view and occlusion-query visibility after a few random steps -- kernels' arithmetic on the CPU vs the oracle."""
    from miniworld_b200.assets import Texture
    from miniworld_b200.envs import LEVELS
    rng = np.random.default_rng(2024)
    for lvl in sorted(LEVELS):
        if lvl in ("MiniWorld-Maze-v0", "MiniWorld-MazeS8-v0"):
            continue          # 8x8 maze: slow on the sequential host sim; MazeS2 / S3 cover the level
        kw = {} if "Sign" in lvl else {"domain_rand": True}
        env = LEVELS[lvl](**kw)
        env.reset(seed=int(rng.integers(0, 10 ** 6)))
        for _ in range(5):
            _, _, te, tr, _ = env.step(int(rng.integers(0, env.action_space.n)))
            if te or tr:
                env.reset()
        obs, depth = env.render_obs(), env.render_depth()
        ts = softgl_lib.TextureSet([t.texels for t in Texture.registry])
        tex_index = lambda tex: tex.tex_id
        rgb, d = softgl_lib.render(env, ts, tex_index)
        assert np.abs(rgb.astype(int) - obs.astype(int)).max() <= 1, lvl
        assert np.array_equal(d, depth), lvl
        top = env.render_top_view()
        assert np.abs(top.astype(int) - softgl_lib.render_top_view(env, ts, tex_index).astype(int)).max() <= 1, lvl
        assert env.get_visible_ents() == softgl_lib.visible_ents(env, ts, tex_index), lvl
        ts.close()
        env.close()


@pytest.mark.parametrize("level,dr", [("MiniWorld-TMaze-v0", True), ("MiniWorld-Sidewalk-v0", False), ("MiniWorld-Sign-v0", False),
                                      ("MiniWorld-CollectHealth-v0", True), ("MiniWorld-PutNext-v0", False),
                                      ("MiniWorld-RoomObjects-v0", True)])
def test_device_programs_equal_python_levels(hostsim_path, level, dr):
    """Seeds and domain_rand settings the golden files do not contain."""
    from helpers import batched_equals_python_levels
    batched_equals_python_levels(level, hostsim_path, dr, n=2, steps=150)


@pytest.mark.parametrize("name,steps", [("hallway", 260), ("fourrooms_dr", 260)])
def test_host_reset_fallback_for_levels_without_a_device_program(hostsim_path, name, steps):
    """A level class that only names its rule (no `device_program`) still runs batched: worlds are generated by its
    Python `_gen_world()` on the host and uploaded (`mwb_set_world`), the env's numpy stream is handed back and forth
    around every host reset (device-side per-step domain-rand draws).  Same reference trajectory, bit for bit."""
    from helpers import state_mismatches
    from miniworld_b200.batched import BatchedMiniWorld
    from miniworld_b200.envs import LEVELS
    level, dr = CASES[name]
    host_only = type("HostOnly" + LEVELS[level].__name__, (LEVELS[level],), {"device_program": None})
    g = golden(name)
    n = 2
    env = BatchedMiniWorld(host_only, n, domain_rand=dr, autoreset=True)
    assert not env.device_reset
    env._host_reset(np.arange(n, dtype=np.int32), [1000 + i for i in range(n)])
    env._seeded = True
    assert not state_mismatches(env, g, 0, n)
    out = None
    for t in range(steps):
        out = env.step_host(g["actions"][t, :n], out, render=False)
        if t % 3 == 2 or t >= 245:           # every state around the truncation at step 250 and the reset after it
            bad = state_mismatches(env, g, t + 1, n, out)
            assert not bad, "step %d: %s" % (t + 1, "; ".join(bad))
    assert g["was_reset"][1:steps + 1, :n].any()
    env.close()


def test_text_frame_with_arbitrary_text(hostsim_path, softgl_lib):
    """reference tests/test_miniworld.py:67-79 (a level that appends a TextFrame with free text), plus the frame
    against the oracle from a pose that looks at the text."""
    import math
    from miniworld_b200.assets import Texture
    from miniworld_b200.entity import TextFrame
    from miniworld_b200.envs import ThreeRooms

    class TestText(ThreeRooms):
        def _gen_world(self):
            super()._gen_world()
            self.entities.append(TextFrame(pos=[0, 1.35, 7], dir=math.pi / 2, str="this is a test"))

    env = TestText()
    env.reset(seed=2)
    env.agent.pos = np.array([0.0, 0.0, 4.0])
    env.agent.dir = -math.pi / 2          # facing +z, towards the wall that carries the text
    obs = env.render_obs()
    ts = softgl_lib.TextureSet([t.texels for t in Texture.registry])
    rgb, _ = softgl_lib.render(env, ts, lambda tex: tex.tex_id)
    assert np.abs(rgb.astype(int) - obs.astype(int)).max() <= 1 and 0 < obs.mean() < 255
    ts.close()
    env.close()


def test_env_checker_invariants(hostsim_path):
    """What gymnasium's check_env verifies for the reference (tests/test_miniworld.py:138-154), restated: seeded
    resets are deterministic, observations live in the declared space, step() returns the five-tuple with the
    documented types."""
    from miniworld_b200.envs import LEVELS
    for eid, cls in LEVELS.items():
        if "Maze-v0" in eid or "MazeS8" in eid:
            continue
        env = cls()
        o1, info1 = env.reset(seed=123)
        p1 = env.agent.pos.copy()
        o2, _ = env.reset(seed=123)
        img = lambda o: o["obs"] if isinstance(o, dict) else o
        assert np.array_equal(img(o1), img(o2)) and np.array_equal(p1, env.agent.pos), eid
        assert isinstance(info1, dict)
        space = env.observation_space
        if isinstance(o1, dict):
            assert set(o1) == {"obs", "goal"} and o1["obs"].shape == (60, 80, 3) and o1["obs"].dtype == np.uint8
        else:
            assert o1.shape == space.shape and o1.dtype == space.dtype, eid
        o, r, te, tr, info = env.step(env.action_space.sample())
        assert isinstance(te, (bool, np.bool_)) and isinstance(tr, (bool, np.bool_)) and isinstance(info, dict), eid
        assert np.isscalar(r) or isinstance(r, (int, float)), eid
        env.close()
