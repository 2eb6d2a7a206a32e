"""The oracle's CPU step port and pixel oracle, pinned: physics against the goldens dumped
from the reference; pixels against small committed oracle frames (regression pin only --
the reference itself offers no golden images, see DESIGN.md 'pixel spec')."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden
from helpers import CASES


@pytest.mark.parametrize("name", ["hallway", "fourrooms", "fourrooms_dr", "pickup"])
def test_step_port_matches_reference_golden(name):
    from miniworld_b200.envs import LEVELS
    from oracle.physics_port import PortEnv
    level, dr = CASES[name]
    g = golden(name)
    for i in range(3):
        port = PortEnv(LEVELS[level](device=None, domain_rand=dr))
        port.reset(seed=1000 + i)
        done = False
        for t in range(150):
            if done:
                port.reset()
                r, te, tr = 0, False, False
            else:
                r, te, tr, removed = port.step(int(g["actions"][t, i]))
                port.finish_pickup(removed)
            done = te or tr
            e = port.env
            assert np.array_equal(e.agent.pos, g["pos"][t + 1, i]) and e.agent.dir == g["dir"][t + 1, i]
            assert r == g["reward"][t + 1, i] and te == g["terminated"][t + 1, i] and tr == g["truncated"][t + 1, i]


def test_pixel_oracle_regression_frames(softgl_lib):
    from miniworld_b200.assets import Texture
    from miniworld_b200.envs import LEVELS
    path = os.path.join(GOLDEN, "oracle_frames.npz")
    frames = {}
    cases = [("MiniWorld-FourRooms-v0", False, 1000), ("MiniWorld-FourRooms-v0", True, 1003),
             ("MiniWorld-Hallway-v0", False, 1001), ("MiniWorld-OneRoom-v0", True, 1002)]
    for level, dr, seed in cases:
        env = LEVELS[level](device=None, domain_rand=dr)
        env.reset(seed=seed)
        ts = softgl_lib.TextureSet([t.texels for t in Texture.registry])
        rgb, depth = softgl_lib.render(env, ts, lambda tex: tex.tex_id)
        ts.close()
        frames["%s_%d_%d_rgb" % (level, dr, seed)] = rgb
        frames["%s_%d_%d_depth" % (level, dr, seed)] = depth
        assert 0 < rgb.mean() < 255 and depth.min() > 0.04 and depth.max() <= 100.0
    if not os.path.exists(path):
        np.savez_compressed(path, **frames)
        pytest.skip("wrote %s" % path)
    pinned = np.load(path)
    for k, v in frames.items():
        assert np.array_equal(v, pinned[k]), k
