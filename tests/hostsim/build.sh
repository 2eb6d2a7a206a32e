#!/bin/sh
# Test-only host simulator: the kernels' MWB_DEV functions compiled for the CPU so that
# kernel logic can be exercised where no GPU exists.  NOT part of the product: the package
# never loads this library; GPU tests, smoke() and bench.py use libmwb.so only.
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
OUT="$HERE/../_hostsim"
mkdir -p "$OUT"
g++ -x c++ -std=c++17 -O2 -ffp-contract=off -mfma -fPIC -shared -DMWB_HOSTSIM \
    -Wno-unused-function -o "$OUT/libmwb_hostsim.so" "$HERE/../../miniworld_b200/csrc/mwb.cu" "$HERE/extra.cpp" -lm
echo "$OUT/libmwb_hostsim.so"
