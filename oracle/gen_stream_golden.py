#!/usr/bin/env python
"""TEST INFRASTRUCTURE (oracle).  Dump golden FRAMES from the UNMODIFIED reference's recorded GL stream.

Replays the golden trajectories of tests/golden/traj_*.npz (same seeds, same actions, same next-step auto-reset
protocol as oracle/gen_golden.py) on /root/reference's own code running under the recording fixed-function GL of
oracle/gl_record.py, and stores what the reference's step() / reset() / render_depth() / render_top_view() /
get_visible_ents() RETURN at selected (step, env) pairs -- the rasterised GL stream -- plus Agent.cam_pos / cam_dir
(entity.py:476-503) at those moments.  These fixtures are what the CUDA rasteriser is compared against on the GPU box
(tests/test_gpu_stream.py) and the CPU host-sim in tests/test_stream_oracle.py.

    python oracle/gen_stream_golden.py [case ...]      # writes tests/golden/stream_*.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
from oracle import ref_stub  # noqa: E402

GOLDEN = os.path.join(HERE, "..", "tests", "golden")

# name -> (reference id, kwargs, envs to record, steps (rows of the trajectory) to record, obs size)
STEPS = (0, 1, 2, 3, 5, 8, 13, 21, 34, 55, 89, 144, 199)
CASES = {
    "hallway": ("MiniWorld-Hallway-v0", {}, (0, 1), STEPS),
    "oneroom": ("MiniWorld-OneRoom-v0", {}, (0, 1), STEPS),
    "fourrooms": ("MiniWorld-FourRooms-v0", {}, (0, 1, 2), STEPS),
    "fourrooms_dr": ("MiniWorld-FourRooms-v0", {"domain_rand": True}, (0, 1, 2), STEPS),
    "pickup": ("MiniWorld-PickupObjects-v0", {}, (0, 1), STEPS),
    "pickup_dr": ("MiniWorld-PickupObjects-v0", {"domain_rand": True}, (0, 1), STEPS),
    "maze_dr": ("MiniWorld-Maze-v0", {"domain_rand": True}, (0, 1, 2), STEPS),
    "mazes3": ("MiniWorld-MazeS3-v0", {}, (0, 1), STEPS),
    "tmaze": ("MiniWorld-TMaze-v0", {}, (0,), STEPS),
    "ymaze_dr": ("MiniWorld-YMaze-v0", {"domain_rand": True}, (0,), STEPS),
    "roomobjs": ("MiniWorld-RoomObjects-v0", {}, (0, 1), STEPS),
    "putnext_dr": ("MiniWorld-PutNext-v0", {"domain_rand": True}, (0, 1), STEPS),
    "wallgap": ("MiniWorld-WallGap-v0", {}, (0,), STEPS),
    "sidewalk_dr": ("MiniWorld-Sidewalk-v0", {"domain_rand": True}, (0,), STEPS),
    "collecthealth": ("MiniWorld-CollectHealth-v0", {}, (0,), STEPS),
    "collecthealth_pick": ("MiniWorld-CollectHealth-v0", {}, (0, 1), STEPS),
    "threerooms_dr": ("MiniWorld-ThreeRooms-v0", {"domain_rand": True}, (0,), STEPS),
    "sign": ("MiniWorld-Sign-v0", {}, (0, 1), (0, 1, 2, 3, 5, 8, 13, 21, 34, 55, 89)),
}
# the same trajectories rendered at BASELINE.json config 5's observation size
# extra envs replayed only for their first steps: known pick-up events of the golden trajectories (env: last row)
EVENT_ENVS = {"pickup": {19: 2, 9: 4, 13: 4}, "pickup_dr": {13: 5, 9: 10}, "pickup_160": {19: 2, 9: 4}}
BIG = {"pickup_160": ("pickup", {"obs_width": 160, "obs_height": 120}, (0,), (0, 3, 21, 89))}


def run_case(name, traj_name, env_id, kwargs, envs, steps):
    g = np.load(os.path.join(GOLDEN, "traj_%s.npz" % traj_name))
    actions = g["actions"]
    T = actions.shape[0]
    env = ref_stub.make_reference_env(env_id, record=True, **kwargs)
    rec = ref_stub.recorder
    steps = sorted(s for s in steps if s <= T)
    sel, rgb, depth, cam_pos, cam_dir, cam_fov, lookat, top, vis, n_tris = [], [], [], [], [], [], [], [], [], []
    events = []

    def grab(t, i, obs, event=False):
        """Record row t of env i.  `obs` is what the reference itself returned for this step."""
        if isinstance(obs, dict):              # Sign's dict observation (sign.py:176)
            obs = obs["obs"]
        assert np.array_equal(env.agent.pos, g["pos"][t, i]) and env.agent.dir == g["dir"][t, i], \
            "trajectory diverged from tests/golden/traj_%s.npz at row %d env %d" % (traj_name, t, i)
        assert len(env.entities) == g["n_ents"][t, i]
        sel.append((t, i))
        rgb.append(obs.copy())
        fr = rec.frames[-1]
        n_tris.append(fr.num_tris())
        lookat.append(fr.view[1:])
        cam_pos.append(np.array(env.agent.cam_pos))
        cam_dir.append(np.array(env.agent.cam_dir))
        cam_fov.append(float(env.agent.cam_fov_y))
        if not event:
            depth.append(env.render_depth())
            assert np.array_equal(env.render_obs(), obs), "render_obs() is not what step() returned"
        else:
            depth.append(np.zeros(obs.shape[:2] + (1,), np.float32))     # a re-render would no longer show the object
        if not event:
            # the map view and the occlusion queries see the entity list AFTER the level's step() edited it
            top.append(env.render_top_view())
            v = 0
            for ent in env.get_visible_ents():
                v |= 1 << env.entities.index(ent)
            vis.append(v)
        else:
            top.append(np.zeros_like(obs))
            vis.append(-1)
        events.append(event)

    extra = EVENT_ENVS.get(name, {})
    for i in list(envs) + [e for e in extra if e not in envs]:
        only_events = i not in envs
        obs, _ = env.reset(seed=1000 + i)
        if 0 in steps and not only_events:
            grab(0, i, obs)
        done = False
        for t in range(extra[i] if only_events else max(steps)):
            if done:
                obs, _ = env.reset()
                te = tr = False
            else:
                obs, r, te, tr, _ = env.step(int(actions[t, i]))
                assert r == g["reward"][t + 1, i]
            done = bool(te or tr)
            # frames in which the base step() drew an object that the level's step() then removed from the entity
            # list (PickupObjects / CollectHealth: the picked-up object is still visible in the returned frame)
            # (detected as: a re-render of the post-step state no longer reproduces the returned frame)
            o = obs["obs"] if isinstance(obs, dict) else obs
            removed = not g["was_reset"][t + 1, i] and not np.array_equal(env.render_obs(), o)
            if (t + 1) in steps and not only_events:
                grab(t + 1, i, obs, event=bool(removed))
            elif removed and sum(events) < 6:
                grab(t + 1, i, obs, event=True)
    out = dict(sel=np.array(sel, np.int32), rgb=np.stack(rgb), depth=np.stack(depth), cam_pos=np.stack(cam_pos),
               cam_dir=np.stack(cam_dir), cam_fov_y=np.array(cam_fov), lookat=np.array(lookat), top=np.stack(top),
               vis=np.array(vis, np.int64), event=np.array(events, bool), n_tris=np.array(n_tris, np.int32),
               meta=np.array([env_id, repr(sorted(kwargs.items())), traj_name, np.__version__]))
    path = os.path.join(GOLDEN, "stream_%s.npz" % name)
    np.savez_compressed(path, **out)
    print("%-20s frames %3d  (%d with a just-removed object)  -> %s (%d B)" % (
        name, len(sel), int(np.sum(events)), os.path.relpath(path), os.path.getsize(path)))


if __name__ == "__main__":
    only = set(sys.argv[1:])
    for name, (env_id, kwargs, envs, steps) in CASES.items():
        if not only or name in only:
            run_case(name, name, env_id, kwargs, envs, steps)
    for name, (traj, kw, envs, steps) in BIG.items():
        if not only or name in only:
            env_id, kwargs = CASES[traj][0], dict(CASES[traj][1], **kw)
            run_case(name, traj, env_id, kwargs, envs, steps)
