#!/bin/bash
# round-2 iteration: parity (stream goldens + oracle) then device-resident K2 timing per MWB_K2_FLAGS setting
# usage: bash tools/gpu_r2.sh "<flags list>" [pytest -k expr]
mkdir -p gpurun_out
FLAGS=${1:-"3 1 0"}
timeout 900 python -m pytest tests/test_gpu_stream.py tests/test_gpu_render.py tests/test_gpu_physics.py -m gpu -q -x ${2:+-k "$2"} > gpurun_out/pytest_r2.log 2>&1
echo "pytest rc=$? $(tail -1 gpurun_out/pytest_r2.log)"
grep -E "FAILED|Error|assert" gpurun_out/pytest_r2.log | head -10
for f in $FLAGS; do
  MWB_K2_FLAGS=$f timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu > gpurun_out/bench_f$f.json 2> gpurun_out/bench_f$f.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_f$f.json"))
    print("flags=$f value=%.0f e2e=%.0f k2_ms=%.4f k1_ms=%.4f clocks=%s" % (d["value"], d["e2e"]["value"], d["roofline"]["kernel_avg_ms"], d["roofline"]["k1_avg_ms"], d["clocks"]))
except Exception as e:
    print("flags=$f bench failed", e); print(open("gpurun_out/bench_f$f.err").read()[-800:])
PY
done
