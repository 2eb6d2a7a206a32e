#!/usr/bin/env python
"""Throughput of every BASELINE.json config at its per-GPU size on one B200 (device-resident
arm only; bench.py is the contract benchmark).  Prints one JSON line per config."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

CONFIGS = [
    # name, level, envs per GPU, kwargs
    ("config1_hallway", "MiniWorld-Hallway-v0", 1, {}),
    ("config2_oneroom", "MiniWorld-OneRoom-v0", 1024, {}),
    ("config3_fourrooms_depth", "MiniWorld-FourRooms-v0", 4096, {"want_depth": True}),
    ("config4_maze_dr_per_gpu", "MiniWorld-MazeS8-v0", 1024, {"domain_rand": True}),
    ("config5_pickup_160x120_per_gpu", "MiniWorld-PickupObjects-v0", 512, {"obs_width": 160, "obs_height": 120}),
]


def main():
    import torch
    from miniworld_b200.batched import BatchedMiniWorld
    steps, warm = 100, 10
    for name, level, n, kw in CONFIGS:
        env = BatchedMiniWorld(level, n, **kw)
        env.reset(seed=1000)
        acts = torch.as_tensor(np.random.default_rng(12345).integers(0, env.action_space.n, size=(steps + warm, n),
                                                                     dtype=np.int32), device="cuda")
        for t in range(warm):
            env.step(acts[t])
        torch.cuda.synchronize()
        env.engine.profile(True)
        env.engine.profile_read()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        done0 = int(env.get_state()["episodes_done"][0])
        e0.record()
        for t in range(warm, warm + steps):
            env.step(acts[t])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        k1, k2, n1, n2 = env.engine.profile_read()
        done = int(env.get_state()["episodes_done"][0]) - done0
        print(json.dumps({"config": name, "level": level, "envs": n, "kwargs": kw, "device_reset": env.device_reset,
                          "env_steps_per_s": n * steps / (ms * 1e-3), "ms_per_step": ms / steps,
                          "k1_ms": k1 / max(1, n1), "k2_ms": k2 / max(1, n2), "episodes_done": done,
                          "tri_overflow_frames": env.engine.overflow_count()}))
        env.close()


if __name__ == "__main__":
    main()
