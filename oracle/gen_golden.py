#!/usr/bin/env python
"""TEST INFRASTRUCTURE (oracle).  Dump golden trajectories from the UNMODIFIED reference.

Runs /root/reference's own miniworld code (GL stubbed by oracle/ref_stub.py) and records,
per environment and step, everything the physics / reward / reset path defines:
agent pos + dir, reward, terminated, truncated, step_count, every entity's pose and the
per-episode domain-randomised parameters.  These are the fixtures under tests/golden/
that the CUDA engine is compared against bit-for-bit.

Protocol (the one BASELINE.md section 4 / SURVEY 8d describe): env i is first reset with
seed = 1000 + i; actions = default_rng(12345).integers(0, n_actions, (T, N)); "next-step"
auto-reset: the step after a terminated|truncated step performs an unseeded reset()
instead of step() (reward 0, flags False), continuing that env's RNG stream.

    python oracle/gen_golden.py            # writes tests/golden/*.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
from oracle.ref_stub import make_reference_env  # noqa: E402

MAX_ENTS = 20

CASES = [
    # name, env id, kwargs, N, T
    ("hallway", "MiniWorld-Hallway-v0", {}, 64, 300),
    ("oneroom", "MiniWorld-OneRoom-v0", {}, 64, 300),
    ("fourrooms", "MiniWorld-FourRooms-v0", {}, 64, 300),
    ("fourrooms_dr", "MiniWorld-FourRooms-v0", {"domain_rand": True}, 32, 300),
    ("pickup", "MiniWorld-PickupObjects-v0", {}, 64, 450),
    ("pickup_dr", "MiniWorld-PickupObjects-v0", {"domain_rand": True}, 16, 300),
    ("maze_dr", "MiniWorld-Maze-v0", {"domain_rand": True}, 64, 300),
    # beyond max_episode_steps = 1536 (maze.py:49: num_rows * num_cols * 24): the truncation branch of step() fires
    ("maze_long", "MiniWorld-Maze-v0", {"domain_rand": True}, 2, 1600),
    ("mazes3", "MiniWorld-MazeS3-v0", {}, 16, 300),
    # levels outside BASELINE.json's configs (single-env path)
    ("tmaze", "MiniWorld-TMaze-v0", {}, 8, 300),
    ("ymaze_dr", "MiniWorld-YMaze-v0", {"domain_rand": True}, 8, 300),
    ("roomobjs", "MiniWorld-RoomObjects-v0", {}, 8, 300),
    ("putnext_dr", "MiniWorld-PutNext-v0", {"domain_rand": True}, 8, 300),
    ("wallgap", "MiniWorld-WallGap-v0", {}, 6, 300),
    ("sidewalk_dr", "MiniWorld-Sidewalk-v0", {"domain_rand": True}, 6, 300),
    ("collecthealth", "MiniWorld-CollectHealth-v0", {}, 4, 300),
    ("threerooms_dr", "MiniWorld-ThreeRooms-v0", {"domain_rand": True}, 6, 300),
    ("sign", "MiniWorld-Sign-v0", {}, 6, 120),
    # pickup-heavy action mix (turns, forward, pickup): exercises CollectHealth's kit respawn
    ("collecthealth_pick", "MiniWorld-CollectHealth-v0", {}, 6, 300, [0.15, 0.15, 0.4, 0.0, 0.3, 0.0, 0.0, 0.0]),
]


def snapshot(env, out, t, i):
    out["pos"][t, i] = env.agent.pos
    out["dir"][t, i] = env.agent.dir
    out["step_count"][t, i] = env.step_count
    out["n_ents"][t, i] = len(env.entities)
    out["carrying"][t, i] = -1
    for e, ent in enumerate(env.entities):
        out["ent_pos"][t, i, e] = ent.pos
        out["ent_dir"][t, i, e] = ent.dir
        out["ent_radius"][t, i, e] = ent.radius
        kind = type(ent).__name__
        out["ent_kind"][t, i, e] = {"Box": 1, "Ball": 2, "Key": 3, "Agent": 4}.get(kind, 9)
        if hasattr(ent, "color_vec"):
            out["ent_color"][t, i, e] = ent.color_vec
        elif hasattr(ent, "mesh"):
            out["ent_color"][t, i, e] = ent.mesh_color if hasattr(ent, "mesh_color") else -1
    out["sky_color"][t, i] = env.sky_color
    out["light_pos"][t, i] = env.light_pos
    out["light_color"][t, i] = env.light_color
    out["light_ambient"][t, i] = env.light_ambient
    a = env.agent
    out["cam"][t, i] = [a.cam_height, a.cam_fwd_disp, a.cam_pitch, a.cam_fov_y]


def run_case(name, env_id, kwargs, N, T, action_probs=None):
    env = make_reference_env(env_id, **kwargs)
    n_act = env.action_space.n
    if action_probs is None:
        actions = np.random.default_rng(12345).integers(0, n_act, size=(T, N), dtype=np.int32)
    else:
        actions = np.random.default_rng(12345).choice(n_act, size=(T, N), p=action_probs).astype(np.int32)
    S = T + 1   # row 0 = state after the seeded reset
    out = dict(
        pos=np.zeros((S, N, 3)), dir=np.zeros((S, N)), step_count=np.zeros((S, N), np.int32),
        reward=np.zeros((S, N)), terminated=np.zeros((S, N), bool), truncated=np.zeros((S, N), bool),
        was_reset=np.zeros((S, N), bool), n_ents=np.zeros((S, N), np.int32),
        carrying=np.zeros((S, N), np.int32),
        ent_pos=np.zeros((S, N, MAX_ENTS, 3)), ent_dir=np.zeros((S, N, MAX_ENTS)),
        ent_radius=np.zeros((S, N, MAX_ENTS)), ent_kind=np.zeros((S, N, MAX_ENTS), np.int8),
        ent_color=np.zeros((S, N, MAX_ENTS, 3)),
        sky_color=np.zeros((S, N, 3)), light_pos=np.zeros((S, N, 3)), light_color=np.zeros((S, N, 3)),
        light_ambient=np.zeros((S, N, 3)), cam=np.zeros((S, N, 4)),
        wall_segs0=None,
    )
    # environments are independent: run them one after the other on one env object
    for i in range(N):
        env.reset(seed=1000 + i)
        if i == 0:
            out["wall_segs0"] = np.array(env.wall_segs)
        snapshot(env, out, 0, i)
        out["was_reset"][0, i] = True
        done = False
        for t in range(T):
            if done:
                env.reset()
                r, te, tr = 0.0, False, False
                out["was_reset"][t + 1, i] = True
            else:
                _, r, te, tr, _ = env.step(int(actions[t, i]))
            done = bool(te or tr)
            out["reward"][t + 1, i] = r
            out["terminated"][t + 1, i] = te
            out["truncated"][t + 1, i] = tr
            snapshot(env, out, t + 1, i)
    out["actions"] = actions
    out["meta"] = np.array([env_id, repr(sorted(kwargs.items())), str(N), str(T), str(n_act),
                            str(env.max_episode_steps), np.__version__])
    path = os.path.join(HERE, "..", "tests", "golden", "traj_%s.npz" % name)
    np.savez_compressed(path, **out)
    print("%-14s N=%d T=%d  episodes ended=%d  reward sum=%.3f  -> %s (%d B)" % (
        name, N, T, int((out["terminated"] | out["truncated"]).sum()), out["reward"].sum(),
        os.path.relpath(path), os.path.getsize(path)))


if __name__ == "__main__":
    only = set(sys.argv[1:])
    for case in CASES:
        if not only or case[0] in only:
            run_case(*case)
