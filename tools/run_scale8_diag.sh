#!/bin/bash
# where does the e2e call lose time with 8 ranks on one host?  traces of rank 0 with and without core pinning
cd "$(dirname "$0")/.."
N=${1:-8}
nvidia-smi topo -m 2>/dev/null | head -14
lscpu | grep -i "numa\|model name\|socket" | head -8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513"
for pin in 0 -1; do
  CV_TRACE=1 timeout 300 $TR bench.py --gpus $N --steps 3 --warmup 3 --e2e-steps 8 --no-cpu-baseline --host-narrow 0 --pin-cores $pin > gpurun_out/diag8_pin$pin.json 2> gpurun_out/diag8_pin$pin.err
  echo "pin=$pin rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/diag8_pin$pin.json")); e=d["e2e"]
print("value %.2f ms  e2e %.2f ms" % (d["ms_per_step"], e["ms_per_step"]), e["ms_each_step"], e["ms_inside_library"])
PY
  grep "cv trace" gpurun_out/diag8_pin$pin.err | tail -16
done
