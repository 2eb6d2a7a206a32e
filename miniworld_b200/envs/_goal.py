"""Shared rule of the "reach the red box" levels."""


class GoalBoxRule:
    """`step()` = base step, then +_reward() and termination once the agent is `near`
    the goal box (reference hallway.py:67-74, oneroom.py:64-71, fourrooms.py:66-73,
    maze.py:155-162).  `device_rule` tells the batched engine to evaluate the same rule
    inside the physics kernel (csrc/physics.cuh, MWB_RULE_GOAL)."""

    device_rule = ("goal", 0)     # entity slot 0 is the box in all four levels

    def step(self, action):
        obs, reward, termination, truncation, info = super().step(action)
        if self.near(self.box):
            reward += self._reward()
            termination = True
        return obs, reward, termination, truncation, info
