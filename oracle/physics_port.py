"""TEST INFRASTRUCTURE (oracle port of the step path).  Not imported by the product.

CPU restatement, in the reference's own numpy/Python idiom, of MiniWorldEnv.step and the
level rules (reference miniworld.py:606-730, 937-975, 1012-1017; envs/hallway.py:67-74;
envs/pickupobjects.py:83-95) on top of the package's host-side world generator.  Uses:
  * a second, independent statement of the physics for CPU tests (validated against the
    reference-generated goldens in tests/test_oracle_port.py);
  * together with oracle/softgl.c, the "port" CPU baseline that bench.py times beside the
    GPU numbers (the reference itself cannot run on the GPU box: no pyglet / GL / gymnasium).
np.linalg.norm on a 3-vector goes through BLAS ddot, whose accumulation (FMA or not) depends
on the CPU; on this image it matches the goldens.
"""
import math

import numpy as np


class PortEnv:
    """Wraps a definition-only env (miniworld_b200.envs.<Level>(device=None)) and steps it on the CPU."""

    def __init__(self, env):
        self.env = env
        rule = getattr(env, "device_rule", None)
        self.rule = rule[0] if rule else None

    def reset(self, seed=None):
        self.env.reset(seed=seed)
        if self.rule == "pickup":
            self.env.num_picked_up = 0

    def _carry_pos(self, agent_pos, ent):
        e = self.env
        dist = e.agent.radius + ent.radius + e.max_forward_step
        pos = agent_pos + e.agent.dir_vec * 1.05 * dist
        y_pos = max(e.agent.cam_height - ent.height - 0.3, 0)
        return pos + np.array([0, 1, 0]) * y_pos

    def _move(self, fwd, drift):
        e = self.env
        a = e.agent
        nxt = a.pos + a.dir_vec * fwd + a.right_vec * drift
        if e.intersect(a, nxt, a.radius):
            return
        if a.carrying:
            cpos = self._carry_pos(nxt, a.carrying)
            if e.intersect(a.carrying, cpos, a.carrying.radius):
                return
            a.carrying.pos = cpos
        a.pos = nxt

    def _turn(self, deg):
        e = self.env
        a = e.agent
        ang = deg * (math.pi / 180)
        old = a.dir
        a.dir += ang
        if a.carrying:
            pos = self._carry_pos(a.pos, a.carrying)
            if e.intersect(a.carrying, pos, a.carrying.radius):
                a.dir = old
                return
            a.carrying.pos = pos
            a.carrying.dir = a.dir

    def step(self, action):
        e = self.env
        a = e.agent
        e.step_count += 1
        rand = e.np_random if e.domain_rand else None
        fwd = e.params.sample(rand, "forward_step")
        drift = e.params.sample(rand, "forward_drift")
        turn = e.params.sample(rand, "turn_step")
        if action == 2:
            self._move(fwd, drift)
        elif action == 3:
            self._move(-fwd, drift)
        elif action == 0:
            self._turn(turn)
        elif action == 1:
            self._turn(-turn)
        elif action == 4:
            test = a.pos + a.dir_vec * 1.5 * a.radius
            hit = e.intersect(a, test, 1.2 * a.radius)
            if not a.carrying and hit is not None and hit is not True and not hit.is_static:
                a.carrying = hit
        elif action == 5:
            if a.carrying:
                a.carrying.pos[1] = 0
                a.carrying = None
        if a.carrying:
            a.carrying.pos = self._carry_pos(a.pos, a.carrying)
            a.carrying.dir = a.dir
        reward, term = 0, False
        trunc = e.step_count >= e.max_episode_steps
        removed = None
        if self.rule == "goal":
            if e.near(e.box):
                reward += e._reward()
                term = True
        elif self.rule == "pickup":
            if a.carrying:
                removed = a.carrying          # still visible in this step's observation
                e.num_picked_up += 1
                reward = 1
                if e.num_picked_up == e.num_objs:
                    term = True
        return reward, term, trunc, removed

    def finish_pickup(self, removed):
        """Apply the post-observation part of the pickup rule (pickupobjects.py:86-88)."""
        if removed is not None:
            self.env.entities.remove(removed)
            self.env.agent.carrying = None
