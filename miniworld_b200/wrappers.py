"""Per-environment Gym wrappers of the reference (miniworld/wrappers.py:7-71): channel-first
observations for PyTorch, greyscale observations, random action replacement."""
import numpy as np

from ._gym import spaces


class _Wrapper:
    """Minimal wrapper base (gymnasium.Wrapper semantics for the calls the tests make)."""

    def __init__(self, env):
        self.env = env
        self.action_space = env.action_space
        self.observation_space = env.observation_space

    def __getattr__(self, name):
        return getattr(self.env, name)

    def reset(self, **kwargs):
        obs, info = self.env.reset(**kwargs)
        return self.observation(obs), info

    def step(self, action):
        obs, reward, terminated, truncated, info = self.env.step(self.action(action))
        return self.observation(obs), reward, terminated, truncated, info

    def observation(self, obs):
        return obs

    def action(self, action):
        return action

    def close(self):
        self.env.close()


class PyTorchObsWrapper(_Wrapper):
    """(H, W, C) -> (C, W, H), the transpose(2, 1, 0) the reference applies."""

    def __init__(self, env):
        super().__init__(env)
        h, w, c = env.observation_space.shape
        self.observation_space = spaces.Box(0, 255, [c, w, h], dtype=env.observation_space.dtype)

    def observation(self, obs):
        return obs.transpose(2, 1, 0)


class GreyscaleWrapper(_Wrapper):
    """RGB -> one luminance channel 0.30 R + 0.59 G + 0.11 B; like the reference (wrappers.py:43-46) the
    result is the float64 array numpy's promotion produces, shape (H, W, 1), not re-quantised."""

    def __init__(self, env):
        super().__init__(env)
        h, w, _ = env.observation_space.shape
        self.observation_space = spaces.Box(0, 255, [h, w, 1], dtype=env.observation_space.dtype)

    def observation(self, obs):
        grey = 0.30 * obs[:, :, 0] + 0.59 * obs[:, :, 1] + 0.11 * obs[:, :, 2]
        return np.expand_dims(grey, axis=2)


class StochasticActionWrapper(_Wrapper):
    """With probability 1 - prob the chosen action is replaced (by `random_action`, else a
    uniformly random one)."""

    def __init__(self, env, prob=0.9, random_action=None):
        super().__init__(env)
        self.prob = prob
        self.random_action = random_action

    def action(self, action):
        if self.env.np_random.uniform() < self.prob:
            return action
        if self.random_action is None:
            return int(self.env.np_random.integers(0, 6))
        return self.random_action
