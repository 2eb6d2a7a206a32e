"""Host-side world generation (miniworld_b200.world / envs) against the reference:
committed golden fixtures always, the live reference too where /root/reference exists."""
import numpy as np
import pytest

from miniworld_b200.envs import LEVELS
from conftest import golden
from helpers import CASES


@pytest.mark.parametrize("name", sorted(CASES))
def test_reset_matches_reference_golden(name):
    level, dr = CASES[name]
    g = golden(name)
    env = LEVELS[level](device=None, **({"domain_rand": True} if dr else {}))    # (Sign fixes domain_rand itself)
    n = min(g["pos"].shape[1], 4 if "maze_dr" in name else 12)
    for i in range(n):
        env.reset(seed=1000 + i)
        assert np.array_equal(env.agent.pos, g["pos"][0, i])
        assert env.agent.dir == g["dir"][0, i]
        assert len(env.entities) == g["n_ents"][0, i]
        for e, ent in enumerate(env.entities):
            assert np.array_equal(np.asarray(ent.pos, float), g["ent_pos"][0, i, e])
            assert float(ent.radius) == g["ent_radius"][0, i, e]
        assert np.array_equal(env.sky_color, g["sky_color"][0, i])
        assert np.array_equal(env.light_pos, g["light_pos"][0, i])
        a = env.agent
        assert [a.cam_height, a.cam_fwd_disp, a.cam_pitch, a.cam_fov_y] == list(g["cam"][0, i])
        if i == 0:
            assert np.array_equal(np.asarray(env.wall_segs), g["wall_segs0"])


def test_live_reference_world_generation():
    from oracle import ref_stub
    if not ref_stub.reference_available():
        pytest.skip("/root/reference not present on this box")
    for eid, kw in [("MiniWorld-FourRooms-v0", {}), ("MiniWorld-FourRooms-v0", {"domain_rand": True}),
                    ("MiniWorld-PickupObjects-v0", {"domain_rand": True}), ("MiniWorld-Hallway-v0", {})]:
        ref = ref_stub.make_reference_env(eid, **kw)
        mine = LEVELS[eid](device=None, **kw)
        for seed in (5, 6, 7):
            ref.reset(seed=seed)
            mine.reset(seed=seed)
            for a, b in zip(ref.entities, mine.entities):
                assert np.array_equal(np.asarray(a.pos, float), np.asarray(b.pos, float)) and a.dir == b.dir
                assert a.radius == b.radius and type(a.radius) is type(b.radius)
            assert np.array_equal(ref.wall_segs, mine.wall_segs)
            for ra, rb in zip(ref.rooms, mine.rooms):
                assert np.array_equal(ra.wall_verts, rb.wall_verts) and np.array_equal(ra.wall_texcs, rb.wall_texcs)
                assert np.array_equal(ra.floor_texcs, rb.floor_texcs) and np.array_equal(ra.wall_norms, rb.wall_norms)
            assert ref.np_random.random() == mine.np_random.random()


def test_reference_level_file_runs_on_this_engine_api():
    """Drop-in check: the reference's own envs/fourrooms.py source, imported against this
    package's MiniWorldEnv / Box, generates the identical world."""
    import importlib.util
    import os
    import sys
    import types
    path = "/root/reference/miniworld/envs/fourrooms.py"
    if not os.path.exists(path):
        pytest.skip("/root/reference not present on this box")
    import miniworld_b200
    from miniworld_b200 import _gym, entity, world
    saved = {k: sys.modules.get(k) for k in ("miniworld", "miniworld.entity", "miniworld.miniworld", "gymnasium")}
    if saved["miniworld"] is not None:
        pytest.skip("the real `miniworld` package is imported in this process")
    try:
        pkg = types.ModuleType("miniworld")
        sys.modules.update({"miniworld": pkg, "miniworld.entity": entity, "miniworld.miniworld": world})
        if not _gym.HAVE_GYMNASIUM:
            shim = types.ModuleType("gymnasium")
            shim.spaces, shim.utils = _gym.spaces, _gym.utils
            sys.modules["gymnasium"] = shim
        spec = importlib.util.spec_from_file_location("ref_fourrooms", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        theirs = mod.FourRooms(device=None)
        ours = miniworld_b200.envs.FourRooms(device=None)
        theirs.reset(seed=3)
        ours.reset(seed=3)
        assert np.array_equal(theirs.agent.pos, ours.agent.pos) and theirs.agent.dir == ours.agent.dir
        assert np.array_equal(theirs.box.pos, ours.box.pos)
        assert np.array_equal(theirs.wall_segs, ours.wall_segs)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


@pytest.mark.parametrize("name", ["fourrooms", "hallway", "mazes3"])
def test_host_move_and_turn_follow_reference(name):
    """MiniWorldEnv.move_agent / turn_agent (host-side mirrors of miniworld.py:620-668, for level code and scripts)
    reproduce the reference trajectory bit for bit up to the first episode end."""
    level, dr = CASES[name]
    g = golden(name)
    env = LEVELS[level](device=None)
    fwd = env.params.sample(None, "forward_step")
    drift = env.params.sample(None, "forward_drift")
    turn = env.params.sample(None, "turn_step")
    for i in range(4):
        env.reset(seed=1000 + i)
        moved = 0
        for t in range(g["actions"].shape[0]):
            if g["terminated"][t, i] or g["truncated"][t, i]:
                break
            a = int(g["actions"][t, i])
            if a == 0:
                env.turn_agent(turn)
            elif a == 1:
                env.turn_agent(-turn)
            elif a == 2:
                moved += bool(env.move_agent(fwd, drift))
            assert np.array_equal(env.agent.pos, g["pos"][t + 1, i]) and env.agent.dir == g["dir"][t + 1, i], (name, i, t)
        assert moved > 0


def test_levels_pickle_like_the_reference():
    """reference tests/test_miniworld.py:157-171: every level survives pickle (EzPickle: constructor arguments) and
    the copy generates the same world from the same seed."""
    import pickle
    for eid, cls in LEVELS.items():
        if "Maze-v0" in eid or "MazeS8" in eid:
            continue
        env = cls(device=None)
        twin = pickle.loads(pickle.dumps(env))
        assert type(twin) is type(env) and twin.max_episode_steps == env.max_episode_steps
        env.reset(seed=5)
        twin.reset(seed=5)
        assert np.array_equal(env.agent.pos, twin.agent.pos) and env.agent.dir == twin.agent.dir, eid
        assert len(env.entities) == len(twin.entities)
