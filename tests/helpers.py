"""Shared comparison code for the CPU (host-sim) and GPU parity tests."""
import numpy as np

from miniworld_b200.batched import BatchedMiniWorld
from miniworld_b200.engine import RNG_DTYPE, rng_state_of

CASES = {
    # golden name: (level id, domain_rand)
    "hallway": ("MiniWorld-Hallway-v0", False),
    "oneroom": ("MiniWorld-OneRoom-v0", False),
    "fourrooms": ("MiniWorld-FourRooms-v0", False),
    "fourrooms_dr": ("MiniWorld-FourRooms-v0", True),
    "pickup": ("MiniWorld-PickupObjects-v0", False),
    "pickup_dr": ("MiniWorld-PickupObjects-v0", True),
    "maze_dr": ("MiniWorld-MazeS8-v0", True),
    "maze_long": ("MiniWorld-MazeS8-v0", True),          # 1600 steps: the max_episode_steps = 1536 truncation fires
    "mazes3": ("MiniWorld-MazeS3-v0", False),
    # levels beyond BASELINE.json's configs whose _gen_world() / rule are lowered too
    "tmaze": ("MiniWorld-TMaze-v0", False),
    "ymaze_dr": ("MiniWorld-YMaze-v0", True),
    "wallgap": ("MiniWorld-WallGap-v0", False),
    "sidewalk_dr": ("MiniWorld-Sidewalk-v0", True),
    "threerooms_dr": ("MiniWorld-ThreeRooms-v0", True),
    "roomobjs": ("MiniWorld-RoomObjects-v0", False),
    "sign": ("MiniWorld-Sign-v0", False),
    "collecthealth": ("MiniWorld-CollectHealth-v0", False),
    "collecthealth_pick": ("MiniWorld-CollectHealth-v0", False),     # pickup-heavy action mix: kit respawns
    "putnext_dr": ("MiniWorld-PutNext-v0", True),
}


def make_env(name, g, lib_path=None, n=None, **kw):
    level, dr = CASES[name]
    N = g["actions"].shape[1] if n is None else n
    env = BatchedMiniWorld(level, N, domain_rand=dr, autoreset=True, **kw)
    seeds = [1000 + i for i in range(N)]
    if env.device_reset:
        env.engine.seed(np.arange(N), np.array([rng_state_of(s) for s in seeds], RNG_DTYPE))
        env.engine.reset()
    else:
        env._host_reset(np.arange(N, dtype=np.int32), seeds)
    env._seeded = True
    return env


def state_mismatches(env, g, t, N, out=None):
    """List of human-readable differences between the engine state and golden row t."""
    st = env.get_state()
    bad = []

    def cmp(label, a, b):
        if not np.array_equal(a, b):
            rows = np.nonzero(~np.all((np.asarray(a) == np.asarray(b)).reshape(len(a), -1), axis=1))[0]
            bad.append("%s: envs %s e.g. got %r want %r" % (label, rows[:4], np.asarray(a)[rows[0]], np.asarray(b)[rows[0]]))

    cmp("agent.pos", st["agent_pos"], g["pos"][t, :N])
    cmp("agent.dir", st["agent_dir"], g["dir"][t, :N])
    cmp("step_count", st["step_count"], g["step_count"][t, :N])
    cmp("cam", st["cam"], g["cam"][t, :N])
    cmp("sky_color", st["env_params"][:, 0:3], g["sky_color"][t, :N])
    cmp("light_pos", st["env_params"][:, 3:6], g["light_pos"][t, :N])
    cmp("light_color", st["env_params"][:, 6:9], g["light_color"][t, :N])
    cmp("light_ambient", st["env_params"][:, 9:12], g["light_ambient"][t, :N])
    ents = st["ents"]
    live = ents["proto"] >= 0
    if not np.array_equal(live.sum(1), g["n_ents"][t, :N]):
        bad.append("entity count: got %r want %r" % (live.sum(1)[:8], g["n_ents"][t, :8]))
    else:
        for i in range(N):
            idx = np.nonzero(live[i])[0]
            if not (np.array_equal(ents[i, idx]["pos"], g["ent_pos"][t, i, :len(idx)]) and
                    np.array_equal(ents[i, idx]["dir"], g["ent_dir"][t, i, :len(idx)])):
                bad.append("entity poses of env %d" % i)
                break
            boxes = g["ent_kind"][t, i, :len(idx)] == 1
            if not np.array_equal(ents[i, idx]["color"][boxes], g["ent_color"][t, i, :len(idx)][boxes]):
                bad.append("box colours of env %d" % i)
                break
    if out is not None:
        cmp("reward", out["reward"], g["reward"][t, :N])
        cmp("terminated", out["terminated"].astype(bool), g["terminated"][t, :N])
        cmp("truncated", out["truncated"].astype(bool), g["truncated"][t, :N])
    return bad


def run_trajectory(name, g, lib_path=None, steps=None, n=None, check_every=1):
    """Replay the golden action stream and compare every state bit-for-bit."""
    env = make_env(name, g, lib_path, n)
    N = env.num_envs
    T = g["actions"].shape[0] if steps is None else min(steps, g["actions"].shape[0])
    bad = state_mismatches(env, g, 0, N)
    assert not bad, "after reset: " + "; ".join(bad)
    out = None
    for t in range(T):
        out = env.step_host(g["actions"][t, :N], out, render=False)
        if (t + 1) % check_every == 0 or t == T - 1:
            bad = state_mismatches(env, g, t + 1, N, out)
            assert not bad, "step %d: %s" % (t + 1, "; ".join(bad))
        else:
            assert np.array_equal(out["reward"], g["reward"][t + 1, :N]), "reward at step %d" % (t + 1)
    env.close()
    return T, N


SINGLE_CASES = {
    # golden name: (level id, kwargs)
    "tmaze": ("MiniWorld-TMaze-v0", {}),
    "ymaze_dr": ("MiniWorld-YMaze-v0", {"domain_rand": True}),
    "roomobjs": ("MiniWorld-RoomObjects-v0", {}),
    "putnext_dr": ("MiniWorld-PutNext-v0", {"domain_rand": True}),
    "wallgap": ("MiniWorld-WallGap-v0", {}),
    "sidewalk_dr": ("MiniWorld-Sidewalk-v0", {"domain_rand": True}),
    "collecthealth": ("MiniWorld-CollectHealth-v0", {}),
    "collecthealth_pick": ("MiniWorld-CollectHealth-v0", {}),
    "threerooms_dr": ("MiniWorld-ThreeRooms-v0", {"domain_rand": True}),
    "sign": ("MiniWorld-Sign-v0", {}),
    "fourrooms": ("MiniWorld-FourRooms-v0", {}),
    "pickup": ("MiniWorld-PickupObjects-v0", {}),
}


def run_single_env_trajectory(name, g, lib_path=None, envs=2, steps=120, device="cuda"):
    """world.MiniWorldEnv (N = 1 engine, host RNG, the level's Python step() rule) against the
    reference trajectory: pose, reward, flags and every entity pose, every step."""
    from miniworld_b200.envs import LEVELS
    level, kw = SINGLE_CASES[name]
    for i in range(envs):
        env = LEVELS[level](device=device, **kw)
        env.reset(seed=1000 + i)
        done = False
        for t in range(min(steps, g["actions"].shape[0])):
            if done:
                env.reset()
                r, te, tr = 0.0, False, False
            else:
                obs, r, te, tr, _ = env.step(int(g["actions"][t, i]))
            done = te or tr
            assert np.array_equal(env.agent.pos, g["pos"][t + 1, i]), (name, i, t)
            assert env.agent.dir == g["dir"][t + 1, i], (name, i, t)
            assert r == g["reward"][t + 1, i] and te == g["terminated"][t + 1, i] and tr == g["truncated"][t + 1, i], (name, i, t)
            assert len(env.entities) == g["n_ents"][t + 1, i]
            for e, ent in enumerate(env.entities):
                assert np.array_equal(np.asarray(ent.pos, float), g["ent_pos"][t + 1, i, e]), (name, i, t, e)
        if isinstance(obs, dict):
            obs = obs["obs"]
        assert obs.shape == (60, 80, 3) and 0 < obs.mean() < 255
        env.close()


def noise_parity(lib_path, n=4, steps=80, prob=0.6, random_action=None):
    """Device-side StochasticActionWrapper == the wrapper around the drop-in single env, draw for draw."""
    from miniworld_b200.batched import BatchedMiniWorld
    from miniworld_b200.engine import RNG_DTYPE, rng_state_of
    from miniworld_b200.envs import Hallway
    from miniworld_b200.wrappers import StochasticActionWrapper
    env = BatchedMiniWorld("MiniWorld-Hallway-v0", num_envs=n, domain_rand=True, autoreset=False)
    ids = np.arange(n, dtype=np.int32)
    env.engine.seed(ids, np.array([rng_state_of(1000 + i) for i in range(n)], RNG_DTYPE))
    env.engine.reset(None)
    env.set_action_noise(prob, random_action)
    singles = [StochasticActionWrapper(Hallway(domain_rand=True), prob=prob, random_action=random_action)
               for _ in range(n)]
    for i, s in enumerate(singles):
        s.reset(seed=1000 + i)
    acts = np.random.default_rng(7).integers(0, 3, size=(steps, n), dtype=np.int32)
    alive = np.ones(n, bool)
    replaced = 0
    for t in range(steps):
        out = env.step_host(acts[t], render=False)
        st = env.get_state()
        for i, s in enumerate(singles):
            if not alive[i]:
                continue
            _, r, te, tr, _ = s.step(int(acts[t, i]))
            a = s.env.agent
            assert np.array_equal(st["agent_pos"][i], a.pos) and st["agent_dir"][i] == a.dir, (t, i)
            assert out["reward"][i] == r and bool(out["terminated"][i]) == te and bool(out["truncated"][i]) == tr
            alive[i] = not (te or tr)
        if not alive.any():
            break
    for s in singles:
        s.close()
    env.close()


def snapshot_roundtrip(level, lib_path, n=6, domain_rand=True, before=25, after=40, **kw):
    """step -> snapshot -> step (recorded) -> restore -> step again: identical; also into a fresh handle."""
    def make():
        e = BatchedMiniWorld(level, num_envs=n, domain_rand=domain_rand, autoreset=True, **kw)
        return e
    env = make()
    ids = np.arange(n, dtype=np.int32)
    env.engine.seed(ids, np.array([rng_state_of(2000 + i) for i in range(n)], RNG_DTYPE))
    env.engine.reset(None)
    nact = env.action_space.n
    acts = np.random.default_rng(3).integers(0, nact, size=(before + after, n), dtype=np.int32)
    for t in range(before):
        env.step_host(acts[t], render=False)
    blob = env.snapshot()

    def run(e):
        rec = []
        for t in range(before, before + after):
            o = e.step_host(acts[t], render=False)
            st = e.get_state(rng=True)
            rec.append((o["reward"].copy(), o["terminated"].copy(), o["truncated"].copy(), st["agent_pos"].copy(),
                        st["agent_dir"].copy(), st["step_count"].copy(), st["rng"].copy()))
        return rec
    first = run(env)
    env.restore(blob)
    second = run(env)
    fresh = make()
    fresh.restore(blob)
    third = run(fresh)
    for a, b, c in zip(first, second, third):
        for x, y, z in zip(a, b, c):
            assert np.array_equal(x, y) and np.array_equal(x, z)
    fresh.close()
    env.close()


def obs_format_parity(lib_path, n=5, steps=6, level="MiniWorld-FourRooms-v0", **kw):
    """K2's fused PyTorchObsWrapper / GreyscaleWrapper epilogues == the wrappers applied to the HWC frames."""
    frames = {}
    for fmt in ("hwc", "cwh", "grey"):
        env = BatchedMiniWorld(level, num_envs=n, autoreset=True, obs_format=fmt, **kw)
        ids = np.arange(n, dtype=np.int32)
        env.engine.seed(ids, np.array([rng_state_of(3000 + i) for i in range(n)], RNG_DTYPE))
        env.engine.reset(None)
        acts = np.random.default_rng(5).integers(0, 3, size=(steps, n), dtype=np.int32)
        out = None
        for t in range(steps):
            out = env.step_host(acts[t], out=out)
        frames[fmt] = out["obs"].copy()
        env.close()
    hwc = frames["hwc"]
    assert frames["cwh"].dtype == np.uint8 and frames["grey"].dtype == np.float64
    for i in range(n):
        assert np.array_equal(frames["cwh"][i], hwc[i].transpose(2, 1, 0))
        o = hwc[i]
        grey = 0.30 * o[:, :, 0] + 0.59 * o[:, :, 1] + 0.11 * o[:, :, 2]
        assert np.array_equal(frames["grey"][i], np.expand_dims(grey, axis=2))
    assert 0 < hwc.mean() < 255


def batched_equals_single_env(level, lib_path, n=3, steps=4, domain_rand=False, **kw):
    """The batched engine (device-side reset program, lowered rule) and the drop-in single-env class (host world
    generation, Python rule) produce identical frames, rewards and flags from the same seeds and actions."""
    from miniworld_b200.envs import LEVELS
    env = BatchedMiniWorld(level, num_envs=n, domain_rand=domain_rand, autoreset=False, **kw)
    assert env.device_reset
    ids = np.arange(n, dtype=np.int32)
    env.engine.seed(ids, np.array([rng_state_of(4000 + i) for i in range(n)], RNG_DTYPE))
    env.engine.reset(None)
    first = np.zeros((n, env.obs_height, env.obs_width, 3), np.uint8)
    env.engine.render(obs=first)
    dr = {"domain_rand": True} if domain_rand else {}
    singles = [LEVELS[level](**dr, **kw) for _ in range(n)]
    for i, s in enumerate(singles):
        o, _ = s.reset(seed=4000 + i)
        o = o["obs"] if isinstance(o, dict) else o
        assert np.array_equal(o, first[i]), (level, "reset frame", i)
    acts = np.random.default_rng(11).integers(0, env.action_space.n, size=(steps, n), dtype=np.int32)
    out = None
    for t in range(steps):
        out = env.step_host(acts[t], out)
        for i, s in enumerate(singles):
            o, r, te, tr, _ = s.step(int(acts[t, i]))
            o = o["obs"] if isinstance(o, dict) else o
            assert np.array_equal(o, out["obs"][i]), (level, "frame", t, i)
            assert r == out["reward"][i] and te == bool(out["terminated"][i]) and tr == bool(out["truncated"][i])
    for s in singles:
        s.close()
    env.close()


def batched_equals_python_levels(level, lib_path, domain_rand, n=3, steps=200, seed0=777):
    """Device-side resets + lowered rule (batched engine) vs the level's own Python `_gen_world()` / `step()` on the
    drop-in class, same seeds and actions, with next-step auto-reset: poses, rewards, flags and every entity position
    are identical at every step.  Independent of the golden files: any seed, either domain_rand setting."""
    from miniworld_b200.envs import LEVELS
    env = BatchedMiniWorld(level, num_envs=n, domain_rand=domain_rand, autoreset=True)
    assert env.device_reset
    ids = np.arange(n, dtype=np.int32)
    env.engine.seed(ids, np.array([rng_state_of(seed0 + i) for i in range(n)], RNG_DTYPE))
    env.engine.reset(None)
    kw = {"domain_rand": True} if domain_rand else {}
    singles = [LEVELS[level](**kw) for _ in range(n)]
    for i, s in enumerate(singles):
        s.reset(seed=seed0 + i)
    na = env.action_space.n
    rng = np.random.default_rng(5)
    p = np.ones(na)
    p[2] = 4                      # forward-heavy (and pickup-heavy) so that goals are reached and things get carried
    if na > 4:
        p[4] = 2
    p /= p.sum()
    done = np.zeros(n, bool)
    episodes, out = 0, None
    for t in range(steps):
        a = rng.choice(na, size=n, p=p).astype(np.int32)
        out = env.step_host(a, out, render=False)
        st = env.get_state()
        for i, s in enumerate(singles):
            if done[i]:
                s.reset()
                r, te, tr = 0.0, False, False
                episodes += 1
            else:
                _, r, te, tr, _ = s.step(int(a[i]))
            done[i] = te or tr
            assert np.array_equal(st["agent_pos"][i], s.agent.pos) and st["agent_dir"][i] == s.agent.dir, (level, t, i)
            assert out["reward"][i] == r and bool(out["terminated"][i]) == te and bool(out["truncated"][i]) == tr, (level, t, i)
            live = [e for e in range(st["ents"].shape[1]) if st["ents"][i, e]["proto"] >= 0]
            assert len(live) == len(s.entities), (level, t, i)
            for e, ent in zip(live, s.entities):
                assert np.array_equal(st["ents"][i, e]["pos"], np.asarray(ent.pos, float)), (level, t, i, e)
    for s in singles:
        s.close()
    env.close()
    return episodes


def stream_cases():
    import glob
    import os
    from conftest import GOLDEN
    return sorted(os.path.basename(p)[len("stream_"):-len(".npz")] for p in glob.glob(os.path.join(GOLDEN, "stream_*.npz")))


def stream_parity(name, lib_path=None, max_rows=None, check_views=True):
    """Frames the UNMODIFIED reference returned while its GL stream was recorded and rasterised
    (tests/golden/stream_*.npz, oracle/gen_stream_golden.py) vs the engine replaying the same trajectory:
    RGB within 1 LSB, depth bit-identical, map view within 1 LSB, occlusion-query sets equal.  Includes the frames in
    which the picked-up object is still drawn although the level's step() already removed it from the entity list."""
    import os
    from conftest import GOLDEN, golden
    with np.load(os.path.join(GOLDEN, "stream_%s.npz" % name)) as z:
        s = {k: z[k] for k in z.files}
    traj = str(s["meta"][2])
    g = golden(traj)
    H, W = s["rgb"].shape[1:3]
    sel = s["sel"]
    if max_rows is not None:
        sel = sel[sel[:, 0] <= max_rows]
    N = int(sel[:, 1].max()) + 1
    env = make_env(traj, g, lib_path, n=N, want_depth=True, obs_width=W, obs_height=H)
    rows = {}
    for k, (t, i) in enumerate(s["sel"]):
        if max_rows is None or t <= max_rows:
            rows.setdefault(int(t), []).append((k, int(i)))
    stats = dict(frames=0, worst=0, same=0, total=0, events=0, tops=0, vis=0, cams=0, cam_exact=0)
    aspect = W / float(H)

    def check_camera(t):
        """K2's camera (mwb_debug_camera) vs the reference's Agent.cam_pos / cam_dir / cam_fov_y at that moment
        (entity.py:476-503) pushed through GLU's gluLookAt / gluPerspective arithmetic in float64 and rounded once:
        every component within 1 float32 ulp (the two float64 routes differ in their last bits), nearly all identical."""
        cam = env.engine.debug_camera()
        for k, i in rows[t]:
            eye, d, fov = s["cam_pos"][k], s["cam_dir"][k], float(s["cam_fov_y"][k])
            assert np.array_equal(s["lookat"][k][:3], eye) and np.array_equal(s["lookat"][k][3:6], eye + d)
            f = (eye + d) - eye
            f = f / np.sqrt(f @ f)
            sv = np.cross(f, [0.0, 1.0, 0.0])
            sv = sv / np.sqrt(sv @ sv)
            u = np.cross(sv, f)
            cot = np.cos(fov / 2 * np.pi / 180) / np.sin(fov / 2 * np.pi / 180)
            want = np.concatenate([eye, sv, u, f, [cot / aspect, cot]]).astype(np.float32)
            got = cam[i, :14]
            ulp = np.spacing(np.maximum(np.abs(want), np.float32(1e-3)))
            assert (np.abs(got - want) <= ulp).all(), "%s row %d env %d: camera %r != %r" % (name, t, i, got, want)
            stats["cams"] += 14
            stats["cam_exact"] += int((got == want).sum())

    out = dict(obs=np.zeros((N, H, W, 3), np.uint8), reward=np.zeros(N), terminated=np.zeros(N, np.uint8),
               truncated=np.zeros(N, np.uint8), depth=np.zeros((N, H, W, 1), np.float32))
    top = np.zeros((N, H, W, 3), np.uint8)
    vis = np.zeros(N, np.int32)

    def check(t):
        top_done = False
        for k, i in rows[t]:
            d = np.abs(out["obs"][i].astype(int) - s["rgb"][k].astype(int))
            stats["frames"] += 1
            stats["worst"] = max(stats["worst"], int(d.max()))
            stats["same"] += int((d == 0).sum())
            stats["total"] += d.size
            assert d.max() <= 1, "%s row %d env %d: %d channel values differ by > 1 LSB (max %d)" % (
                name, t, i, int((d > 1).sum()), int(d.max()))
            if s["event"][k]:
                stats["events"] += 1
                continue
            assert np.array_equal(out["depth"][i], s["depth"][k]), "%s row %d env %d: depth differs" % (name, t, i)
            if not check_views:
                continue
            if not top_done:
                env.render_top_view(out=top)
                env.visible_ents(out=vis)
                top_done = True
            dt = np.abs(top[i].astype(int) - s["top"][k].astype(int))
            assert dt.max() <= 1, "%s row %d env %d: top view differs by %d" % (name, t, i, int(dt.max()))
            stats["tops"] += 1
            # device bits are per entity-list SLOT; the golden's are per list index
            st = env.get_state()["ents"][i]
            live = [e for e in range(len(st)) if st[e]["proto"] >= 0]
            got = sum(1 << j for j, e in enumerate(live) if (int(vis[i]) >> e) & 1)
            assert got == int(s["vis"][k]), "%s row %d env %d: visible set %s != %s" % (name, t, i, bin(got), bin(int(s["vis"][k])))
            stats["vis"] += 1

    if 0 in rows:
        env.engine.render(obs=out["obs"], depth=out["depth"])
        check(0)
        check_camera(0)
    for t in range(1, max(rows) + 1):
        need = t in rows
        env.step_host(g["actions"][t - 1, :N], out, render=need)
        if need:
            check(t)
            check_camera(t)
    env.close()
    return stats
