#!/bin/bash
# Compare K2 launch variants (MWB_K2_VARIANT) on one B200: render parity tests + device-resident bench each.
mkdir -p gpurun_out
for v in ${VARIANTS:-0 1 2}; do
  export MWB_K2_VARIANT=$v
  MWB_DEBUG=1 timeout 600 python -m pytest tests/test_gpu_render.py -m gpu -q -x > gpurun_out/pytest_v$v.log 2>&1; echo "variant $v pytest rc=$? $(grep -E 'passed|failed' gpurun_out/pytest_v$v.log | tail -1) $(grep -m1 '\[mwb\]' gpurun_out/pytest_v$v.log)"
  MWB_DEBUG=1 timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu > gpurun_out/bench_v$v.json 2> gpurun_out/bench_v$v.err; grep -m1 '\[mwb\]' gpurun_out/bench_v$v.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_v$v.json"))
print("variant $v value=%.0f k2_ms=%.3f k1_ms=%.3f" % (d["value"], d["roofline"]["kernel_avg_ms"], d["roofline"]["k1_avg_ms"]))
PY
done
