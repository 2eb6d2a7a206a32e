#!/bin/bash
# ncu capture of K2 + mesh pre-pass on the PickupObjects 160x120 config.
mkdir -p gpurun_out
cat > /tmp/pk.py <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, '.')
from miniworld_b200.batched import BatchedMiniWorld
env = BatchedMiniWorld("MiniWorld-PickupObjects-v0", 512, obs_width=160, obs_height=120)
env.reset(seed=1000)
acts = torch.as_tensor(np.random.default_rng(1).integers(0, 5, size=(8, 512), dtype=np.int32), device="cuda")
for t in range(8):
    env.step(acts[t])
torch.cuda.synchronize()
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"render_kernel|mesh_setup" -s 8 -c 2 -f -o gpurun_out/prof_pickup python /tmp/pk.py > gpurun_out/ncu_pickup.log 2>&1; echo "ncu rc=$?"
