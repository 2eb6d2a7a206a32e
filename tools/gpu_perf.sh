#!/bin/bash
# Perf iteration on one B200: GPU tests (subset via $1), bench, one full ncu capture of K2.
mkdir -p gpurun_out
KEXPR=${1:-"render or device_reset or host_reset or properties or pickup"}
timeout 1500 python -m pytest tests -m gpu -q -x -k "$KEXPR" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
grep -E "passed|failed|worst|Error" gpurun_out/pytest_gpu.log | tail -8
timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench.json"))
print("value=%.0f e2e=%.0f k2_ms=%.3f k1_ms=%.3f cpu=%s frac=%.4f" % (d["value"], d["e2e"]["value"], d["roofline"]["kernel_avg_ms"], d["roofline"]["k1_avg_ms"], d["cpu_baseline"]["value"], d["roofline"]["frac"]))
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 4 -c 1 -f -o gpurun_out/prof_k2 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1; echo "ncu rc=$?"
timeout 600 python tools/bench_configs.py > gpurun_out/configs.jsonl 2> gpurun_out/configs.err; echo "configs rc=$?"; cat gpurun_out/configs.jsonl | cut -c1-400; tail -3 gpurun_out/configs.err
