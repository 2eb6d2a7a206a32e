"""BatchedMiniWorld: N independent MiniWorld environments stepped by one C-ABI call.

This is the hot path of the package: `step(actions)` = K1 (physics / reward / auto-reset,
one warp per env) + K2 (first-person render, one block per env) through `mwb_step`.
Per-environment semantics are exactly those of the reference's `MiniWorldEnv.reset/step`
(miniworld.py:544-604, 670-730) plus the level's own `step()` rule, for every env:

  * env i seeded with `reset(seed=[...])` owns the numpy stream
    Generator(PCG64(SeedSequence(seed_i))), continued across unseeded resets;
  * levels with a fixed room layout (Hallway, OneRoom, FourRooms, PickupObjects) reset
    entirely on the device from the lowered `device_program` (csrc/reset.cuh);
  * levels whose topology is random per episode (Maze) generate worlds with the level's
    Python `_gen_world()` on the host (same numpy stream) and upload them (`mwb_set_world`);
  * `autoreset=True` gives Gymnasium "next-step" auto-reset: the step after a
    terminated|truncated step resets that env (action ignored, reward 0).

Outputs are torch CUDA tensors by default (zero-copy from the kernels); `step_host`
performs the same step with pinned host buffers for callers that want numpy.
"""
import numpy as np

from . import pack
from .engine import Engine, RULE_GOAL, RULE_NONE, RULE_HEALTH, RULE_PICKUP, RULE_PUTNEXT, RULE_SIDEWALK, RULE_SIGN, generator_from_state, rng_state_of, RNG_DTYPE
from .envs import LEVELS
from .program import ResetProgram


def _resolve_level(level):
    if isinstance(level, str):
        if level not in LEVELS:
            raise KeyError("unknown level id %r (known: %s)" % (level, ", ".join(sorted(LEVELS))))
        return LEVELS[level]
    return level


def _torch_stream(torch, device):
    """Handle of torch's current stream for mwb_step / mwb_render_obs.  torch's default stream is the legacy default
    stream (handle 0); a NULL stream argument means "the handle's own stream" to the C ABI, which would not be ordered
    after the torch kernels that produced the actions -- so the default stream is passed as cudaStreamLegacy (0x1)."""
    return torch.cuda.current_stream(device).cuda_stream or 1


class BatchedMiniWorld:
    def __init__(self, level, num_envs, obs_width=80, obs_height=60, domain_rand=False, autoreset=True,
                 msaa_samples=8, device=0, want_depth=False, level_kwargs=None, obs_format="hwc"):
        self.level_cls = _resolve_level(level)
        self.level_kwargs = dict(level_kwargs or {})
        self.num_envs = int(num_envs)
        self.obs_width, self.obs_height = int(obs_width), int(obs_height)
        self.domain_rand = bool(domain_rand)
        self.want_depth = bool(want_depth)
        self.device = int(device)

        # a definition-only instance of the level: layout, params, rule, action space
        dr = {"domain_rand": True} if self.domain_rand else {}     # (Sign fixes domain_rand itself, like the reference)
        self.proto_env = self.level_cls(device=None, obs_width=obs_width, obs_height=obs_height, **dr,
                                        **self.level_kwargs)
        pe = self.proto_env
        self.action_space = pe.action_space                  # per-env space: `step` takes one action per env
        self.single_action_space = pe.action_space           # (gymnasium.vector naming)
        self.single_observation_space = pe.observation_space
        self.max_episode_steps = pe.max_episode_steps
        rule = getattr(pe, "device_rule", None)
        if rule is None:
            raise TypeError("%s has no `device_rule`; use world.MiniWorldEnv (single env) for levels whose "
                            "step() rule is not lowered" % self.level_cls.__name__)
        rule = {"goal": RULE_GOAL, "pickup": RULE_PICKUP, "sidewalk": RULE_SIDEWALK, "sign": RULE_SIGN, "health": RULE_HEALTH, "putnext": RULE_PUTNEXT, "none": RULE_NONE}[rule[0]], rule[1]
        self.device_reset = getattr(pe, "device_program", None) is not None

        rooms, quads, segs = pack.pack_geometry(pe)
        self.maze_template = None
        if self.device_reset:
            self.program = ResetProgram()
            pe.device_program(self.program)
            if self.program.uses_maze:
                # per-episode topology on the device: only if the translated templates reproduce
                # host-generated worlds exactly; otherwise fall back to host-side resets
                from .maze_lowering import MazeTemplate
                try:
                    tmpl = MazeTemplate(self.level_cls, domain_rand=self.domain_rand, **self.level_kwargs)
                    tmpl.verify(seeds=(0,))
                    self.maze_template = tmpl
                except AssertionError:
                    self.device_reset = False
        if self.device_reset:
            protos = self.program.proto_array()
            max_ents = max(2, self.program.num_placed)     # exact: K2's shared-memory triangle capacity scales with it
            caps = (len(rooms), len(quads), len(segs))
            if self.maze_template is not None:
                caps = (len(rooms), len(quads) + 8, len(segs) + 8)
        else:
            self.program = None
            protos = None
            max_ents = 8
            caps = tuple(int(1.25 * n) + 4 for n in (len(rooms), len(quads), len(segs)))
        shared = self.device_reset and self.maze_template is None
        self.engine = Engine(self.num_envs, obs_width, obs_height, msaa_samples,
                             shared_geometry=shared, max_rooms=caps[0], max_quads=caps[1],
                             max_segs=caps[2], max_ents=max_ents, rule=rule, domain_rand=self.domain_rand,
                             max_episode_steps=int(min(self.max_episode_steps, 2 ** 31 - 1)),   # math.inf: never truncates
                             autoreset=autoreset and self.device_reset,
                             device=device)
        self.autoreset = bool(autoreset)
        eng = self.engine
        eng.sync_assets()
        eng.set_params(pe.params)
        if self.device_reset:
            eng.set_protos(protos)
            if self.maze_template is not None:
                eng.set_maze(self.maze_template, pack.room_cdf(pe.room_probs))
            else:
                eng.set_template(rooms, quads, segs)
            eng.set_program(self.program.op_array())
        else:
            # host-reset levels: one worker env per slot keeps that env's RNG stream
            self._workers = [None] * self.num_envs
            self._host_done = np.zeros(self.num_envs, bool)
        # observation layout written by the render kernel: the reference's PyTorchObsWrapper ("cwh") and
        # GreyscaleWrapper ("grey") are fused into its epilogue instead of running as separate passes
        self.obs_format = obs_format
        N, H, W = self.num_envs, self.obs_height, self.obs_width
        self.obs_shape = {"hwc": (N, H, W, 3), "cwh": (N, 3, W, H), "grey": (N, H, W, 1)}[obs_format]
        self.obs_dtype = np.float64 if obs_format == "grey" else np.uint8
        if obs_format != "hwc":
            eng.set_obs_format(obs_format)
        self._seeded = False
        self._torch = None
        self._bufs = None

    # ------------------------------------------------------------------ buffers
    def _ensure_torch(self):
        if self._torch is None:
            import torch
            self._torch = torch
            dev = torch.device("cuda", self.device)
            N, H, W = self.num_envs, self.obs_height, self.obs_width
            self._bufs = dict(
                obs=torch.zeros(self.obs_shape, dtype=torch.float64 if self.obs_format == "grey" else torch.uint8, device=dev),
                depth=torch.zeros((N, H, W, 1), dtype=torch.float32, device=dev) if self.want_depth else None,
                reward=torch.zeros(N, dtype=torch.float64, device=dev),
                terminated=torch.zeros(N, dtype=torch.uint8, device=dev),
                truncated=torch.zeros(N, dtype=torch.uint8, device=dev),
                actions=torch.zeros(N, dtype=torch.int32, device=dev),
            )
            # bool views of the uint8 flag buffers the kernel writes (no per-step conversion kernels)
            self._bufs["term_view"] = self._bufs["terminated"].view(torch.bool)
            self._bufs["trunc_view"] = self._bufs["truncated"].view(torch.bool)
            self._info = {"depth": self._bufs["depth"]} if self.want_depth else {}
            # what the level's step() puts into `info` / the observation, read in place from the device state
            eng = self.engine
            self._info_views = {}
            for key, spec in (getattr(self.proto_env, "device_info", None) or {}).items():
                if spec[0] == "counter":                   # CollectHealth: info["health"] (collecthealth.py:100)
                    self._info[key] = torch.as_tensor(eng.state_array("counter"), device=dev)
                elif spec[0] == "entity_pos":              # TMaze: info["goal_pos"] = self.box.pos (tmaze.py:89)
                    self._info_views[key] = (int(spec[1]), [torch.as_tensor(eng.state_array(n), device=dev)
                                                            for n in ("ent_x", "ent_y", "ent_z")])
            self._obs_dict = dict(getattr(self.proto_env, "device_obs_extra", None) or {})
        return self._torch

    def _wrap(self, obs):
        """Observation / info as the level's own step() shapes them (Sign: {"obs", "goal"}, sign.py:176)."""
        torch = self._torch
        for key, (slot, xyz) in self._info_views.items():
            self._info[key] = torch.stack([a[slot] for a in xyz], dim=1)          # float64 [N, 3]
        if self._obs_dict:
            obs = dict({k: torch.full((self.num_envs,), int(v), dtype=torch.int64, device=obs.device)
                        for k, v in self._obs_dict.items()}, obs=obs)
        return obs

    # ------------------------------------------------------------------ reset
    def reset(self, seed=None, env_ids=None):
        """Reset all (or the listed) envs.  `seed`: int base (env i gets seed + i), a sequence
        of per-env seeds, or None to continue each env's stream.  Returns (obs, info)."""
        ids = np.arange(self.num_envs, dtype=np.int32) if env_ids is None else np.asarray(env_ids, np.int32)
        if seed is not None:
            seeds = [int(seed) + int(i) for i in ids] if np.isscalar(seed) else [int(s) for s in seed]
            assert len(seeds) == len(ids)
        else:
            seeds = None
            if not self._seeded:
                seeds = [int(s) for s in np.random.SeedSequence().generate_state(len(ids))]
        if self.device_reset:
            if seeds is not None:
                states = np.array([rng_state_of(s) for s in seeds], RNG_DTYPE)
                self.engine.seed(ids, states)
            self.engine.reset(None if env_ids is None else ids)
        else:
            self._host_reset(ids, seeds)
        self._seeded = True
        return self._wrap(self.render()), {}

    def _host_reset(self, ids, seeds, hold=False):
        """Host-side reset of the listed envs (levels without a device program).  The env's numpy
        stream is shared with the device: per-step domain-rand draws happen in K1, so the
        stream position is pulled from the device before `_gen_world()` runs on the host and
        pushed back afterwards."""
        worlds = []
        dev_rng = self.engine.get_state(rng=True)["rng"] if (self.domain_rand and self._seeded) else None
        for k, i in enumerate(ids):
            w = self._workers[i]
            if w is None:
                w = self._workers[i] = self.level_cls.__new__(self.level_cls)
                w.__dict__.update({k2: v for k2, v in self.proto_env.__dict__.items()
                                   if k2 not in ("_np_random", "agent", "entities", "rooms", "wall_segs")})
                w._np_random = None
            elif dev_rng is not None and (seeds is None):
                w._np_random = generator_from_state(dev_rng[i])
            w.reset(seed=None if seeds is None else seeds[k])
            worlds.append(pack.pack_world(w))
            worlds[-1]["hold"] = int(hold)
        # one proto table for the handle: per-env protos are identical for these levels
        self.engine.sync_assets()
        self.engine.set_protos(worlds[0]["protos"])
        self.engine.set_world(ids, worlds)
        if self.domain_rand:
            self.engine.seed(ids, np.array([rng_state_of(self._workers[i].np_random) for i in ids], RNG_DTYPE))

    # ------------------------------------------------------------------ step / render
    def step(self, actions):
        """actions: int tensor / array [N].  Returns (obs, reward, terminated, truncated, info)
        as torch CUDA tensors; obs is uint8 [N, H, W, 3] (and info['depth'] when want_depth)."""
        torch = self._ensure_torch()
        b = self._bufs
        if isinstance(actions, torch.Tensor) and actions.is_cuda and actions.dtype == torch.int32 and actions.is_contiguous():
            acts = actions                      # consumed in place by K1, no copy
        elif isinstance(actions, torch.Tensor):
            b["actions"].copy_(actions.to(torch.int32), non_blocking=True)
            acts = b["actions"]
        else:
            b["actions"].copy_(torch.as_tensor(np.asarray(actions, np.int32)), non_blocking=True)
            acts = b["actions"]
        stream = _torch_stream(torch, self.device)
        if not self.device_reset and self.autoreset and self._host_done.any():
            self._host_reset(np.nonzero(self._host_done)[0].astype(np.int32), None, hold=True)
            self._host_done[:] = False
        self.engine.step(acts, obs=b["obs"], depth=b["depth"], reward=b["reward"],
                         terminated=b["terminated"], truncated=b["truncated"], stream=stream)
        if not self.device_reset and self.autoreset:
            self._host_done = (b["terminated"] | b["truncated"]).bool().cpu().numpy()
        return self._wrap(b["obs"]), b["reward"], b["term_view"], b["trunc_view"], self._info

    def step_host(self, actions, out=None, render=True):
        """Same step with HOST buffers end to end (numpy in, numpy out): actions are copied
        host->device and obs / reward / flags device->host inside the call."""
        N, H, W = self.num_envs, self.obs_height, self.obs_width
        if out is None:
            out = dict(obs=np.zeros(self.obs_shape, self.obs_dtype), reward=np.zeros(N), terminated=np.zeros(N, np.uint8),
                       truncated=np.zeros(N, np.uint8),
                       depth=np.zeros((N, H, W, 1), np.float32) if self.want_depth else None)
        acts = np.ascontiguousarray(actions, np.int32)
        if not self.device_reset and self.autoreset and self._host_done.any():
            self._host_reset(np.nonzero(self._host_done)[0].astype(np.int32), None, hold=True)
            self._host_done[:] = False
        self.engine.step(acts, obs=out["obs"] if render else None, depth=out.get("depth") if render else None,
                         reward=out["reward"], terminated=out["terminated"], truncated=out["truncated"])
        if not self.device_reset and self.autoreset:
            self._host_done = (out["terminated"] | out["truncated"]).astype(bool)
        return out

    def render(self):
        torch = self._ensure_torch()
        b = self._bufs
        stream = _torch_stream(torch, self.device)
        self.engine.render(obs=b["obs"], depth=b["depth"], stream=stream)
        return b["obs"]

    def render_depth(self):
        torch = self._ensure_torch()
        dev = torch.device("cuda", self.device)
        d = torch.zeros((self.num_envs, self.obs_height, self.obs_width, 1), dtype=torch.float32, device=dev)
        self.engine.render(depth=d, stream=_torch_stream(torch, self.device))
        return d

    def snapshot(self):
        """Checkpoint of all envs (numpy uint8 blob); `restore` resumes them bit for bit."""
        if not self.device_reset:
            raise TypeError("snapshot covers the device-resident state; host-reset levels keep RNG streams in Python")
        return self.engine.snapshot()

    def restore(self, blob):
        self.engine.restore(blob)
        self._seeded = True

    def set_action_noise(self, prob=0.9, random_action=None):
        """Device-side StochasticActionWrapper (reference wrappers.py:49-71) for every env: the replacement draws
        come from each env's own numpy stream, in the order the wrapper would make them.  prob=None disables."""
        if not self.device_reset and prob is not None:
            raise TypeError("action noise needs the env streams on the device (levels with a device reset program)")
        self.engine.set_action_noise(prob, random_action)

    def render_top_view(self, render_agent=True, out=None):
        """Map view of every env in the observation layout -- uint8 [N, H, W, 3] by default (reference
        render_top_view, miniworld.py:1088-1175).
        Extents come from the level definition (all envs of a level share them).  `out`: optional numpy
        array / CUDA tensor to fill; default a fresh CUDA tensor."""
        ext = self.proto_env.top_view_extents(self.obs_width, self.obs_height)
        if out is None:
            torch = self._ensure_torch()
            out = torch.zeros(self.obs_shape, dtype=torch.float64 if self.obs_format == "grey" else torch.uint8,
                              device=torch.device("cuda", self.device))      # same layout as the observations
        self.engine.render_top_view(ext, out, render_agent)
        return out

    def visible_ents(self, out=None):
        """uint32 [N]: bit e set iff entity-list slot e of that env passes the reference's occlusion query
        (get_visible_ents, miniworld.py:1238-1333)."""
        if out is None:
            torch = self._ensure_torch()
            out = torch.zeros(self.num_envs, dtype=torch.int32, device=torch.device("cuda", self.device))
        self.engine.visible_ents(out)
        return out

    # ------------------------------------------------------------------ state views
    def get_state(self, **kw):
        return self.engine.get_state(**kw)

    def np_random(self, i):
        """numpy Generator positioned where env i's device stream currently is."""
        return generator_from_state(self.engine.get_state(rng=True)["rng"][i])

    def close(self):
        self.engine.close()
