"""K1 on a real B200 through the C ABI: bit-exact against trajectories dumped from the
unmodified reference (tests/golden, oracle/gen_golden.py)."""
import numpy as np
import pytest

from conftest import golden
from helpers import CASES, run_trajectory

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["hallway", "oneroom", "fourrooms", "fourrooms_dr", "pickup", "pickup_dr"])
def test_device_reset_levels_bit_exact(libmwb_path, name):
    """pose / dir / reward / terminated / truncated / entity poses / per-episode domain
    randomisation, every step, device-side resets included."""
    T, N = run_trajectory(name, golden(name), libmwb_path, check_every=5)   # reward / flags every step
    assert T >= 300 and N >= 16


@pytest.mark.parametrize("name", ["tmaze", "ymaze_dr", "wallgap", "sidewalk_dr", "threerooms_dr", "roomobjs", "sign", "collecthealth", "collecthealth_pick", "putnext_dr"])
def test_lowered_extra_levels_bit_exact(libmwb_path, name):
    """Levels beyond BASELINE.json's configs on the batched engine (IFEQ / PUT reset ops, street rule)."""
    run_trajectory(name, golden(name), libmwb_path, check_every=5)


@pytest.mark.parametrize("name", ["mazes3", "maze_dr", "maze_long"])
def test_maze_levels_bit_exact(libmwb_path, name):
    """Device-generated mazes (csrc/maze.cuh); maze_dr: BASELINE.json config 4 at N = 64 x T = 300;
    maze_long: 1600 steps, so the truncation at max_episode_steps = 1536 (maze.py:49) and the reset after it happen."""
    g = golden(name)
    T, N = run_trajectory(name, g, libmwb_path, check_every=5)
    if name == "maze_long":
        assert T == 1600 and g["truncated"].sum() >= 1 and g["step_count"].max() == 1536


def test_rng_stream_position_after_rollout(libmwb_path):
    """After 300 steps with resets the device PCG64 state equals numpy's after the same
    draws: replay the host mirror's resets on the side and compare one more draw."""
    from helpers import make_env
    from miniworld_b200.envs import LEVELS
    g = golden("fourrooms")
    env = make_env("fourrooms", g, libmwb_path, n=8)
    out = None
    for t in range(300):
        out = env.step_host(g["actions"][t, :8], out, render=False)
    for i in range(8):
        mirror = LEVELS["MiniWorld-FourRooms-v0"](device=None)
        mirror.reset(seed=1000 + i)
        for _ in range(int(g["was_reset"][1:, i].sum())):
            mirror.reset()
        assert env.np_random(i).random() == mirror.np_random.random()
    env.close()


@pytest.mark.parametrize("level,dr", [("MiniWorld-TMaze-v0", True), ("MiniWorld-YMaze-v0", False), ("MiniWorld-WallGap-v0", True),
                                      ("MiniWorld-Sidewalk-v0", False), ("MiniWorld-ThreeRooms-v0", False),
                                      ("MiniWorld-RoomObjects-v0", True), ("MiniWorld-Sign-v0", False),
                                      ("MiniWorld-CollectHealth-v0", True), ("MiniWorld-PutNext-v0", False),
                                      ("MiniWorld-PickupObjects-v0", True), ("MiniWorld-MazeS3Fast-v0", True)])
def test_device_programs_equal_python_levels(libmwb_path, level, dr):
    """Batched engine (device reset program + lowered rule) vs the level's Python `_gen_world()` / `step()` on the
    drop-in class over many episodes, for seeds / domain_rand settings the golden files do not contain."""
    from helpers import batched_equals_python_levels
    batched_equals_python_levels(level, libmwb_path, dr, n=4, steps=400)


@pytest.mark.parametrize("name", ["hallway", "fourrooms_dr"])
def test_host_reset_fallback_for_levels_without_a_device_program(libmwb_path, name):
    """Levels that only name their rule (user-defined ones): host `_gen_world()` + mwb_set_world, RNG stream handed
    back and forth around every host reset; same reference trajectory bit for bit."""
    from helpers import state_mismatches
    from miniworld_b200.batched import BatchedMiniWorld
    from miniworld_b200.envs import LEVELS
    level, dr = CASES[name]
    host_only = type("HostOnly" + LEVELS[level].__name__, (LEVELS[level],), {"device_program": None})
    g = golden(name)
    n = 8
    env = BatchedMiniWorld(host_only, n, domain_rand=dr, autoreset=True)
    assert not env.device_reset
    env._host_reset(np.arange(n, dtype=np.int32), [1000 + i for i in range(n)])
    env._seeded = True
    out = None
    for t in range(300):
        out = env.step_host(g["actions"][t, :n], out, render=False)
        bad = state_mismatches(env, g, t + 1, n, out)
        assert not bad, "step %d: %s" % (t + 1, "; ".join(bad))
    env.close()
