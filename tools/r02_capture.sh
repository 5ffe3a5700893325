#!/bin/bash
# ncu captures of round 2: launch list of config 4, --set full of the GEMM kernel on configs 3 / 4 / 5-shard
# (run under gpurun; outputs in gpurun_out/)
set -x
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 --clock-sampler off"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02j_launches_cfg4.csv $B --workload cfg4_bulk > gpurun_out/r02j_ncu_cfg4_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_dense_contract -s 2 -c 1 -o gpurun_out/r02j_cfg4_gemm -f $B --workload cfg4_bulk > gpurun_out/r02j_ncu_cfg4.log 2>&1
ncu --set full --clock-control none -k regex:k_dense_rows_gather -s 1 -c 1 -o gpurun_out/r02j_cfg4_rows -f $B --workload cfg4_bulk > gpurun_out/r02j_ncu_cfg4_rows.log 2>&1
ncu --set full --clock-control none -k regex:k_dense_contract -s 3 -c 1 -o gpurun_out/r02j_cfg3_gemm -f $B --workload cfg3_kmers > gpurun_out/r02j_ncu_cfg3.log 2>&1
ncu --set full --clock-control none -k regex:k_dense_contract -s 3 -c 1 -o gpurun_out/r02j_cfg5_gemm -f $B --workload cfg5_atlas_shard > gpurun_out/r02j_ncu_cfg5.log 2>&1
ls -la gpurun_out/
