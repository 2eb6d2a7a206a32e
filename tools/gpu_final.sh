#!/bin/bash
# round-end evidence on one B200: smoke, the whole -m gpu suite, the default bench line, the reference arm, the ncu
# launch list of the bench command and one `ncu --set full` capture of K2 with source counters.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$? $(tail -1 gpurun_out/final_smoke.log)"
timeout 1500 python -m pytest tests -m gpu -q --durations=12 > gpurun_out/final_pytest_gpu.log 2>&1; echo "pytest rc=$? $(grep -E 'passed|failed' gpurun_out/final_pytest_gpu.log | tail -1)"
grep -E "FAILED|Error" gpurun_out/final_pytest_gpu.log | head -10
grep -A14 "slowest" gpurun_out/final_pytest_gpu.log | head -16
timeout 600 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/final_bench.json")); r=d["roofline"]
print("value=%.0f e2e=%.0f k2=%.4f k1=%.4f frac=%.4f cpu=%s launches=%s clocks=%s" % (d["value"], d["e2e"]["value"], r["kernel_avg_ms"], r["k1_avg_ms"], r["frac"], d["cpu_baseline"]["value"], d["gpu_launches"], d["clocks"]))
PY
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/final_bench_reference.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/final_bench_reference.json')); print('reference arm: %.0f env-steps/s on %d cores (%.0f per core)' % (d['value'], d['cpu_baseline']['cores'], d['config']['per_physical_core']))"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu > /dev/null 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 4 -c 1 -f -o gpurun_out/prof_k2_final python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_final.log 2>&1; echo "ncu rc=$?"
