#!/bin/bash
# bench + one full ncu capture of K2 (source-level counters) on one B200
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_render.py -m gpu -q -x > gpurun_out/pytest_prof.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/pytest_prof.log)"
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu > gpurun_out/bench_prof.json 2> gpurun_out/bench_prof.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_prof.json"))
print("value=%.0f e2e=%.0f k2_ms=%.3f k1_ms=%.3f" % (d["value"], d["e2e"]["value"], d["roofline"]["kernel_avg_ms"], d["roofline"]["k1_avg_ms"]))
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 4 -c 1 -f -o gpurun_out/prof_k2 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1; echo "ncu rc=$?"
