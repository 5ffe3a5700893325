#!/bin/bash
# 8-GPU weak-scaling runs (one box): config 2 per GPU, then the atlas (config 5: 500k peaks x 1M cells,
# 125k cells per GPU) when the host has the memory for 8 x 17 GB of pinned buffers.
cd "$(dirname "$0")/.."
N=${1:-8}
echo "cores $(nproc)"; free -g | head -2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_cfg2_${N}gpu.json 2> gpurun_out/bench_cfg2_${N}gpu.err
echo "cfg2 rc=$?"; grep "timed step" gpurun_out/bench_cfg2_${N}gpu.err | head -$N
timeout 300 $TR bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_cfg2_${N}gpu_ref.json 2> gpurun_out/bench_cfg2_${N}gpu_ref.err
echo "reference arm rc=$?"
avail=$(free -g | awk '/^Mem:/ {print $7}')
if [ "$avail" -ge $(( N * 28 )) ]; then
  timeout 900 $TR bench.py --gpus $N --workload cfg5_atlas_shard --steps 3 --warmup 3 > gpurun_out/bench_cfg5_${N}gpu.json 2> gpurun_out/bench_cfg5_${N}gpu.err
  echo "cfg5 rc=$?"; grep "timed step" gpurun_out/bench_cfg5_${N}gpu.err | head -$N
else
  echo "cfg5 skipped: ${avail} GB available"
fi
cat gpurun_out/bench_cfg2_${N}gpu.json gpurun_out/bench_cfg2_${N}gpu_ref.json gpurun_out/bench_cfg5_${N}gpu.json 2>/dev/null | cut -c1-1500
