#!/bin/bash
# 8-GPU evidence run (gpurun --gpus 8): weak / strong scaling and the two 8-GPU configs of BASELINE.json, the end-to-end
# leg with and without NUMA binding, NVLink byte counters of GPU 0 around the weak-scaling run.
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_${N}gpu.txt 2>&1
port=29900
run() {  # tag, extra args...
  tag=$1; shift
  port=$((port+1))
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 60 --warmup 10 --no-cpu "$@" > gpurun_out/multi_${tag}_${N}gpu.json 2> gpurun_out/multi_${tag}_${N}gpu.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/multi_${tag}_${N}gpu.json"))
    r=d["roofline"]
    print("${tag} N=$N value=%.0f ms/step=%.4f k2=%.4f k1=%.4f non_kernel=%.4f e2e=%.0f d2h=%s" % (d["value"], d["ms_per_step"], r["kernel_avg_ms"], r["k1_avg_ms"], r["non_kernel_ms_per_step"], d["e2e"]["value"], ["%.1f"%x for x in d["e2e"]["d2h_gbs_per_rank"]]))
except Exception as e:
    print("${tag} failed:", e); print(open("gpurun_out/multi_${tag}_${N}gpu.err").read()[-1500:])
PY
}
nvidia-smi nvlink -gt d -i 0 > gpurun_out/nvlink_before.txt 2>&1
run c3weak
nvidia-smi nvlink -gt d -i 0 > gpurun_out/nvlink_after.txt 2>&1
run c3strong --scaling strong
run c4 --config 4
run c5 --config 5
# (the --no-numa comparison was taken once: profiles/r2_8gpu_first_run.txt)
python - <<'PY'
import re
def tot(p):
    rx=tx=0
    for l in open(p):
        m=re.search(r"Data (Rx|Tx): (\d+) KiB", l)
        if m:
            if m.group(1)=="Rx": rx+=int(m.group(2))
            else: tx+=int(m.group(2))
    return rx,tx
try:
    b,a=tot("gpurun_out/nvlink_before.txt"),tot("gpurun_out/nvlink_after.txt")
    print("GPU0 NVLink during c3weak (70 steps + warm-up/e2e legs): rx %.1f MB, tx %.1f MB" % ((a[0]-b[0])/1024.0,(a[1]-b[1])/1024.0))
except Exception as e:
    print("nvlink counters unavailable:", e)
PY
