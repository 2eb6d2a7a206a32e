import sys, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
HS = sys.argv[1]
from miniworld_b200 import engine
engine._override_library_for_tests(HS)      # tools-only seam: the product classes take no library argument
from conftest import golden
from helpers import (run_trajectory, noise_parity, snapshot_roundtrip, obs_format_parity, batched_equals_single_env,
                     batched_equals_python_levels)
for name, steps in [("fourrooms_dr", 60), ("pickup_dr", 80), ("mazes3", 60), ("maze_dr", 12), ("collecthealth_pick", 200),
                    ("putnext_dr", 80), ("sign", 60), ("sidewalk_dr", 80), ("threerooms_dr", 60), ("ymaze_dr", 60)]:
    run_trajectory(name, golden(name), HS, steps=steps, n=3, check_every=20); print("traj", name, "ok", flush=True)
noise_parity(HS, n=2, steps=30); print("noise ok", flush=True)
snapshot_roundtrip("MiniWorld-PickupObjects-v0", HS, n=3, before=10, after=15); print("snapshot ok", flush=True)
obs_format_parity(HS, n=2, steps=1); print("formats ok", flush=True)
obs_format_parity(HS, n=2, steps=1, obs_width=44, obs_height=30); print("ragged formats ok", flush=True)
batched_equals_single_env("MiniWorld-PutNext-v0", HS, n=2, steps=2, domain_rand=True); print("frames ok", flush=True)
batched_equals_single_env("MiniWorld-CollectHealth-v0", HS, n=2, steps=2); print("frames2 ok", flush=True)
from miniworld_b200.envs import LEVELS
env = LEVELS["MiniWorld-ThreeRooms-v0"](domain_rand=True); env.reset(seed=1)
env.render_top_view(); env.get_visible_ents(); env.render_depth(); env.close(); print("views ok", flush=True)
print("asan run complete")
