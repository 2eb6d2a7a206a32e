#!/bin/bash
# GPU tests (all) + one full ncu capture of K2.  Logs -> gpurun_out/.
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
grep -E "passed|failed|worst" gpurun_out/pytest_gpu.log | tail -12
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 4 -c 2 -f -o gpurun_out/prof_k2 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/ncu_full.log
