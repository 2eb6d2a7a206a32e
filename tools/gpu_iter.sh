#!/bin/bash
# One development iteration on a B200: selected GPU tests ($1 = pytest -k expression), then the K2 variants.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -k "${1:-top_view or render}" > gpurun_out/pytest_iter.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_iter.log
VARIANTS="${VARIANTS:-1 2}" bash tools/gpu_variants.sh 2>&1 | grep -v "^$"
