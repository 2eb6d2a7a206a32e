#!/bin/bash
# CPU-only: the host simulator (the kernels' MWB_DEV functions + the whole host side of mwb.cu) under
# AddressSanitizer + UBSan, driven through trajectories, resets, respawns, rendering, top view, visibility,
# snapshot / restore and the fused observation layouts.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=${TMPDIR:-/tmp}/mwb_asan
mkdir -p "$OUT"
g++ -x c++ -std=c++17 -O1 -g -ffp-contract=off -mfma -fPIC -shared -DMWB_HOSTSIM -fsanitize=address,undefined \
    -fno-omit-frame-pointer -Wno-unused-function -o "$OUT/libmwb_hostsim_asan.so" \
    "$ROOT/miniworld_b200/csrc/mwb.cu" "$ROOT/tests/hostsim/extra.cpp" -lm
ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 UBSAN_OPTIONS=print_stacktrace=1 \
  LD_PRELOAD=$(gcc -print-file-name=libasan.so) python "$ROOT/tools/asan_hostsim_run.py" "$OUT/libmwb_hostsim_asan.so"
