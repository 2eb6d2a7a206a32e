#!/bin/bash
# mesh-level checks on one B200: parity tests that draw meshes, then the per-config throughput table
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -k "pickup or single_env or all_levels or top_view or render" > gpurun_out/pytest_mesh.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_mesh.log
timeout 900 python tools/bench_configs.py > gpurun_out/configs.jsonl 2> gpurun_out/configs.err; echo "configs rc=$?"
python - <<PY
import json
for l in open("gpurun_out/configs.jsonl"):
    d=json.loads(l); print("%-34s %10.0f steps/s  k1 %.3f ms  k2 %.3f ms  step %.3f ms" % (d["config"], d["env_steps_per_s"], d["k1_ms"], d["k2_ms"], d["ms_per_step"]))
PY
tail -3 gpurun_out/configs.err
