#!/bin/bash
# Perf iteration on one B200: render/physics GPU tests, bench for both K2 register variants, ncu.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -k "render or smoke or device_reset or properties or pickup" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
grep -E "passed|failed|worst|Error" gpurun_out/pytest_gpu.log | tail -8
for mb in 3 2; do
  MWB_K2_MINBLOCKS=$mb timeout 600 python bench.py --steps 60 --warmup 10 --no-cpu > gpurun_out/bench_mb$mb.json 2> gpurun_out/bench_mb$mb.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_mb$mb.json"))
print("minblocks=$mb value=%.0f e2e=%.0f k2_ms=%.3f k1_ms=%.3f" % (d["value"], d["e2e"]["value"], d["roofline"]["kernel_avg_ms"], d["roofline"]["k1_avg_ms"]))
PY
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 4 -c 1 -f -o gpurun_out/prof_k2 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1; echo "ncu rc=$?"
