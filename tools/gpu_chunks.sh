#!/bin/bash
# end-to-end (host buffers) throughput vs number of render/copy pipeline pieces
mkdir -p gpurun_out
for c in ${CHUNKS:-4 8 16 32}; do
  MWB_D2H_CHUNKS=$c timeout 600 python bench.py --steps 60 --warmup 6 --no-cpu > gpurun_out/bench_c$c.json 2> gpurun_out/bench_c$c.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_c$c.json"))
print("chunks $c value=%.0f e2e=%.0f e2e_ms=%.3f k2_ms=%.3f" % (d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["kernel_avg_ms"]))
PY
done
