#!/bin/bash
# 8-GPU A/B of the e2e call with and without host-side narrowing of the counts' values (same box).
cd "$(dirname "$0")/.."
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512"
for hn in 1 0; do
  timeout 400 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --host-narrow $hn > gpurun_out/bench_cfg2_${N}gpu_hn$hn.json 2> gpurun_out/bench_cfg2_${N}gpu_hn$hn.err
  echo "host_narrow=$hn rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_cfg2_${N}gpu_hn$hn.json")); e=d["e2e"]
print("value %.3e (%.2f ms)  e2e %.3e (%.2f ms) h2d %d" % (d["value"], d["ms_per_step"], e["value"], e["ms_per_step"], e["h2d_bytes_per_step"]), e["ms_each_step"])
PY
done
