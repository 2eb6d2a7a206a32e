#!/bin/bash
# multi-GPU iteration (use with gpurun --gpus N): peer-path parity tests, then torchrun benches per config / partition
# usage: bash tools/gpu_multi2.sh N "<runs>"   with runs = list of  tag:extra-bench-args  (spaces inside args as commas)
N=${1:-2}
RUNS=${2:-"c3weak: c3strong:--scaling,strong c4:--config,4 c5:--config,5"}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_properties.py -m gpu -q -x -k "peer or ordered or snapshot_after" > gpurun_out/pytest_multi.log 2>&1
echo "pytest rc=$? $(tail -1 gpurun_out/pytest_multi.log)"; grep -E "FAILED|Error" gpurun_out/pytest_multi.log | head
port=29810
for run in $RUNS; do
  tag=${run%%:*}; extra=$(echo "${run#*:}" | tr ',' ' ')
  port=$((port+1))
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 100 --warmup 10 --no-cpu $extra > gpurun_out/multi_${tag}_${N}gpu.json 2> gpurun_out/multi_${tag}_${N}gpu.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/multi_${tag}_${N}gpu.json"))
    r=d["roofline"]
    print("${tag} N=$N value=%.0f ms/step=%.4f k2=%.4f k1=%.4f non_kernel=%.4f e2e=%.0f d2h=%s | %s" % (d["value"], d["ms_per_step"], r["kernel_avg_ms"], r["k1_avg_ms"], r["non_kernel_ms_per_step"], d["e2e"]["value"], ["%.1f"%x for x in d["e2e"]["d2h_gbs_per_rank"]], d["config"]["obs_gather"][:40]))
except Exception as e:
    print("${tag} failed:", e); print(open("gpurun_out/multi_${tag}_${N}gpu.err").read()[-1500:])
PY
done
