"""K2 on a real B200 against the CPU pixel oracle (oracle/softgl.c): RGB within 1 LSB,
depth bit-identical (=> far inside the 1e-4 relative bound), at reset and along rollouts."""
import numpy as np
import pytest

from conftest import golden
from helpers import CASES, make_env

pytestmark = pytest.mark.gpu


def oracle_frames(softgl, level, dr, env, ids, width=80, height=60):
    """Oracle render of engine envs `ids`: host mirror world + the engine's dynamic state."""
    from miniworld_b200.assets import Texture
    from miniworld_b200.envs import LEVELS
    from miniworld_b200.engine import generator_from_state
    st = env.get_state(room_tex=True)
    ts = softgl.TextureSet([t.texels for t in Texture.registry])
    out = []
    mirror = LEVELS[level](device=None, domain_rand=dr)
    for i in ids:
        m = mirror
        # geometry is static for these levels; copy the dynamic state from the engine
        ents = st["ents"][i]
        live = [e for e in range(len(ents)) if ents[e]["proto"] >= 0]
        assert len(live) == len(m.entities), "mirror and engine disagree on the entity list"
        for e, ent in zip(live, m.entities):
            ent.pos, ent.dir = np.array(ents[e]["pos"]), float(ents[e]["dir"])
            if hasattr(ent, "color_vec"):
                ent.color_vec = np.array(ents[e]["color"])
        a = m.agent
        a.cam_height, a.cam_fwd_disp, a.cam_pitch, a.cam_fov_y = st["cam"][i]
        m.sky_color, m.light_pos = st["env_params"][i, 0:3], st["env_params"][i, 3:6]
        m.light_color, m.light_ambient = st["env_params"][i, 6:9], st["env_params"][i, 9:12]
        tex_of = {}
        for r, room in enumerate(m.rooms):
            tex_of[id(room.wall_tex)] = st["room_tex"][i, r, 0]
        # per-room texture variants chosen on the device
        def tex_index(tex, _m=m, _i=i):
            return tex.tex_id
        if dr:
            # rebuild texcoords for the variants the device picked
            from miniworld_b200.world import gen_texcs_floor, gen_texcs_wall
            for r, room in enumerate(m.rooms):
                wt, ft, ct = (Texture.registry[k] for k in st["room_tex"][i, r])
                room.wall_tex, room.floor_tex, room.ceil_tex = wt, ft, ct
                room.floor_texcs = gen_texcs_floor(ft, room.floor_verts)
                room.ceil_texcs = gen_texcs_floor(ct, room.ceil_verts)
                scale = np.array([512 / wt.width, 512 / wt.height])
                room.wall_texcs = (room.wall_uvm * scale).astype(np.float32)
        out.append(softgl.render(m, ts, tex_index, width, height))
    ts.close()
    return out


@pytest.mark.parametrize("name", ["hallway", "oneroom", "fourrooms", "fourrooms_dr"])
def test_rollout_frames_match_oracle(libmwb_path, softgl_lib, name):
    level, dr = CASES[name]
    g = golden(name)
    n = 16
    env = make_env(name, g, libmwb_path, n=n, want_depth=True)
    out = None
    worst, exact, total = 0, 0, 0
    for t in range(40):
        out = env.step_host(g["actions"][t, :n], out)
        if t % 8 == 7 or t == 0:
            ids = list(range(n)) if t == 0 else [t % n, (3 * t) % n]
            for i, (rgb, depth) in zip(ids, oracle_frames(softgl_lib, level, dr, env, ids)):
                diff = np.abs(rgb.astype(int) - out["obs"][i].astype(int))
                worst = max(worst, int(diff.max()))
                exact += int((diff == 0).sum())
                total += diff.size
                assert diff.max() <= 1, "env %d step %d: %d pixels differ by > 1 LSB" % (i, t, (diff > 1).sum())
                assert np.array_equal(depth, out["depth"][i]), "env %d step %d: depth codes differ" % (i, t)
                assert 0 < out["obs"][i].mean() < 255
    print("%s: worst |diff| = %d LSB, %.4f%% of channel values identical" % (name, worst, 100.0 * exact / total))
    env.close()


@pytest.mark.parametrize("msaa", [1, 4, 8, 16])
def test_sample_counts(libmwb_path, softgl_lib, msaa):
    from miniworld_b200.assets import Texture
    from miniworld_b200.envs import LEVELS
    g = golden("fourrooms")
    env = make_env("fourrooms", g, libmwb_path, n=4, want_depth=True, msaa_samples=msaa)
    obs = env.render().cpu().numpy()
    ts = softgl_lib.TextureSet([t.texels for t in Texture.registry])
    for i in range(4):
        ref = LEVELS["MiniWorld-FourRooms-v0"](device=None)
        ref.reset(seed=1000 + i)
        rgb, _ = softgl_lib.render(ref, ts, lambda tex: tex.tex_id, samples=msaa)
        assert np.abs(rgb.astype(int) - obs[i].astype(int)).max() <= 1
    ts.close()
    env.close()


def test_human_view_is_rendered_with_16_samples(libmwb_path, softgl_lib):
    """render() of the drop-in class = the reference's vis_fb frame: FrameBuffer(window_width, window_height, 16)
    (miniworld.py:518), agent view and map view, against the oracle with the 16-sample pattern (which the reference's own
    render() equals under the recording GL: tests/test_stream_oracle.py)."""
    from miniworld_b200.assets import Texture
    from miniworld_b200.envs import Hallway
    for view in ("agent", "top"):
        env = Hallway(render_mode="rgb_array", window_width=200, window_height=150, view=view)
        env.reset(seed=3)
        env.step(2)
        frame = env.render()
        ts = softgl_lib.TextureSet([t.texels for t in Texture.registry])
        want = (softgl_lib.render(env, ts, lambda tex: tex.tex_id, 200, 150, 16)[0] if view == "agent"
                else softgl_lib.render_top_view(env, ts, lambda tex: tex.tex_id, 200, 150, 16))
        ts.close()
        d = np.abs(frame.astype(int) - want.astype(int))
        assert frame.shape == (150, 200, 3) and d.max() <= 1 and (d == 0).mean() > 0.995
        env.close()


def test_obs_160x120(libmwb_path, softgl_lib):
    from miniworld_b200.assets import Texture
    from miniworld_b200.envs import LEVELS
    g = golden("fourrooms")
    env = make_env("fourrooms", g, libmwb_path, n=4, obs_width=160, obs_height=120)
    obs = env.render().cpu().numpy()
    assert obs.shape == (4, 120, 160, 3)
    ts = softgl_lib.TextureSet([t.texels for t in Texture.registry])
    for i in range(4):
        ref = LEVELS["MiniWorld-FourRooms-v0"](device=None)
        ref.reset(seed=1000 + i)
        rgb, _ = softgl_lib.render(ref, ts, lambda tex: tex.tex_id, 160, 120)
        assert np.abs(rgb.astype(int) - obs[i].astype(int)).max() <= 1
    ts.close()
    env.close()


@pytest.mark.parametrize("w,h,level", [(84, 62, "MiniWorld-FourRooms-v0"), (45, 31, "MiniWorld-PickupObjects-v0")])
def test_ragged_frame_sizes(libmwb_path, softgl_lib, w, h, level):
    """Frame sizes that are not multiples of the 8x4 half-tile: edge tiles, byte-wise stores, and the fused
    channel-first / greyscale epilogues on them."""
    from helpers import obs_format_parity
    from miniworld_b200.assets import Texture
    from miniworld_b200.batched import BatchedMiniWorld
    from miniworld_b200.envs import LEVELS
    env = BatchedMiniWorld(level, 3, obs_width=w, obs_height=h, want_depth=True)
    env.reset(seed=1000)
    obs = env.render().cpu().numpy()
    depth = env.render_depth().cpu().numpy()
    assert obs.shape == (3, h, w, 3)
    ts = softgl_lib.TextureSet([t.texels for t in Texture.registry])
    for i in range(3):
        ref = LEVELS[level](device=None, obs_width=w, obs_height=h)
        ref.reset(seed=1000 + i)
        rgb, d = softgl_lib.render(ref, ts, lambda tex: tex.tex_id, w, h)
        assert np.abs(rgb.astype(int) - obs[i].astype(int)).max() <= 1
        assert np.array_equal(d, depth[i])
    ts.close()
    env.close()
    obs_format_parity(libmwb_path, n=3, steps=2, level=level, obs_width=w, obs_height=h)


def test_pickup_objects_meshes_160x120(libmwb_path, softgl_lib):
    """Ball / Key meshes (5192 / 208 triangles, un-normalised normals) and boxes, 160x120,
    at reset and after steps incl. the frame in which a picked-up object is shown carried."""
    from miniworld_b200.assets import Texture
    from miniworld_b200.envs import LEVELS
    g = golden("pickup")
    n = 8
    env = make_env("pickup", g, libmwb_path, n=n, want_depth=True, obs_width=160, obs_height=120)
    ts = softgl_lib.TextureSet([t.texels for t in Texture.registry])
    out = None
    checked = 0
    for t in range(-1, 60):
        if t >= 0:
            out = env.step_host(g["actions"][t, :n], out)
            obs, depth = out["obs"], out["depth"]
        else:
            obs = env.render().cpu().numpy()
            depth = env.render_depth().cpu().numpy()
        picked = t >= 0 and (out["reward"] == 1).any()
        if t in (-1, 20, 59) or picked:
            st = env.get_state()
            for i in ([int(np.argmax(out["reward"]))] if picked else range(n)):
                if picked:
                    continue     # the carried ("ghost") object is not part of get_state; covered on the host sim
                m = LEVELS["MiniWorld-PickupObjects-v0"](device=None)
                m.reset(seed=1000 + i)
                ents = st["ents"][i]
                live = [e for e in range(len(ents)) if ents[e]["proto"] >= 0]
                if len(live) != len(m.entities):
                    continue     # something was picked up earlier: entity list differs from the fresh mirror
                for e, ent in zip(live, m.entities):
                    ent.pos, ent.dir = np.array(ents[e]["pos"]), float(ents[e]["dir"])
                rgb, d = softgl_lib.render(m, ts, lambda tex: tex.tex_id, 160, 120)
                diff = np.abs(rgb.astype(int) - obs[i].astype(int))
                assert diff.max() <= 1, "env %d step %d: %d values differ by > 1 LSB" % (i, t, (diff > 1).sum())
                assert np.array_equal(d, depth[i])
                checked += 1
    assert checked >= 16
    ts.close()
    env.close()


@pytest.mark.parametrize("name", ["mazes3", "maze_dr"])
def test_maze_frames_match_oracle(libmwb_path, softgl_lib, name):
    """Device-generated mazes (csrc/maze.cuh) rendered by K2 vs the oracle on the host-generated
    world of the same seed; the 8x8 maze keeps its triangle lists in HBM (no overflow allowed)."""
    from miniworld_b200.assets import Texture
    from miniworld_b200.envs import LEVELS
    level, dr = CASES[name]
    g = golden(name)
    n = 6
    env = make_env(name, g, libmwb_path, n=n, want_depth=True)
    assert env.device_reset
    obs = env.render().cpu().numpy()
    depth = env.render_depth().cpu().numpy()
    ts = softgl_lib.TextureSet([t.texels for t in Texture.registry])
    for i in range(n):
        m = LEVELS[level](device=None, domain_rand=dr)
        m.reset(seed=1000 + i)
        rgb, d = softgl_lib.render(m, ts, lambda tex: tex.tex_id)
        diff = np.abs(rgb.astype(int) - obs[i].astype(int))
        assert diff.max() <= 1, "env %d: %d values differ by > 1 LSB" % (i, (diff > 1).sum())
        assert np.array_equal(d, depth[i])
    assert env.engine.overflow_count() == 0
    ts.close()
    env.close()


@pytest.mark.parametrize("level", ["MiniWorld-Hallway-v0", "MiniWorld-PickupObjects-v0", "MiniWorld-ThreeRooms-v0",
                                   "MiniWorld-MazeS3-v0", "MiniWorld-CollectHealth-v0"])
def test_top_view_and_visible_ents_single_env(libmwb_path, softgl_lib, level):
    """render_top_view / get_visible_ents of the drop-in class (reference miniworld.py:1088-1175, 1238-1333)
    vs the immediate-mode oracle."""
    from miniworld_b200.assets import Texture
    from miniworld_b200.envs import LEVELS
    env = LEVELS[level]()
    for seed in (3, 4, 5):
        env.reset(seed=seed)
        for _ in range(6):
            env.step(env.action_space.sample())
        ts = softgl_lib.TextureSet([t.texels for t in Texture.registry])
        top = env.render_top_view()
        ref = softgl_lib.render_top_view(env, ts, lambda tex: tex.tex_id)
        assert np.abs(top.astype(int) - ref.astype(int)).max() <= 1
        assert (top == ref).mean() > 0.999
        assert env.get_visible_ents() == softgl_lib.visible_ents(env, ts, lambda tex: tex.tex_id)
        ts.close()
    env.close()


def test_top_view_and_visible_ents_batched(libmwb_path, softgl_lib):
    """The batched entry points (mwb_render_top_view / mwb_visible_ents over N envs) vs the oracle on host worlds
    generated from the same seeds; both visible and hidden goal boxes must occur."""
    from miniworld_b200.assets import Texture
    from miniworld_b200.batched import BatchedMiniWorld
    from miniworld_b200.envs import LEVELS
    N = 24
    env = BatchedMiniWorld("MiniWorld-FourRooms-v0", num_envs=N)
    env.reset(seed=500)
    tops = env.render_top_view().cpu().numpy()
    masks = env.visible_ents().cpu().numpy().astype(np.uint32)
    ts = softgl_lib.TextureSet([t.texels for t in Texture.registry])
    seen = set()
    for i in range(N):
        ref = LEVELS["MiniWorld-FourRooms-v0"](device=None)
        ref.reset(seed=500 + i)
        want = softgl_lib.render_top_view(ref, ts, lambda tex: tex.tex_id)
        assert np.abs(tops[i].astype(int) - want.astype(int)).max() <= 1
        vis = softgl_lib.visible_ents(ref, ts, lambda tex: tex.tex_id)
        want_mask = sum(1 << e for e, ent in enumerate(ref.entities) if ent in vis)
        assert int(masks[i]) == want_mask, (i, int(masks[i]), want_mask)
        seen.add(want_mask != 0)
    assert seen == {True, False}
    ts.close()
    env.close()


@pytest.mark.parametrize("level,dr", [("MiniWorld-PutNext-v0", True), ("MiniWorld-Sign-v0", False), ("MiniWorld-TMaze-v0", False),
                                      ("MiniWorld-YMaze-v0", True), ("MiniWorld-WallGap-v0", False),
                                      ("MiniWorld-ThreeRooms-v0", True), ("MiniWorld-Sidewalk-v0", True),
                                      ("MiniWorld-RoomObjects-v0", False), ("MiniWorld-CollectHealth-v0", False)])
def test_batched_frames_equal_single_env(libmwb_path, level, dr):
    """Levels lowered beyond BASELINE.json's configs: the batched engine's frames (device reset program, per-episode
    box sizes, fixed-pose meshes, text / image frames) == the drop-in class's, whose frames the oracle tests pin."""
    from helpers import batched_equals_single_env
    batched_equals_single_env(level, libmwb_path, n=6, steps=12, domain_rand=dr)


def test_large_batch_frames_equal_small_batch(libmwb_path):
    """At N >= 1776 a frame is rendered by ONE block and leaves the SM through the whole-frame shared-memory
    stage (16-byte stores); small batches split a frame over several blocks and store row segments directly.
    Same seeds => same worlds => the two routes must give identical frames, in both uint8 layouts."""
    import torch
    from miniworld_b200.batched import BatchedMiniWorld
    small_n, big_n, steps = 48, 2048, 3
    acts = np.random.default_rng(9).integers(0, 3, size=(steps, big_n), dtype=np.int32)
    for fmt in ("hwc", "cwh"):
        frames = {}
        for n in (small_n, big_n):
            env = BatchedMiniWorld("MiniWorld-FourRooms-v0", n, obs_format=fmt, want_depth=True)
            env.reset(seed=1000)
            for t in range(steps):
                obs, _, _, _, info = env.step(torch.as_tensor(acts[t, :n], device="cuda"))
            frames[n] = (obs[:small_n].cpu().numpy().copy(), info["depth"][:small_n].cpu().numpy().copy())
            env.close()
        assert np.array_equal(frames[small_n][0], frames[big_n][0]), fmt
        assert np.array_equal(frames[small_n][1], frames[big_n][1]), fmt
        assert 0 < frames[big_n][0].mean() < 255
