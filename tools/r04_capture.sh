#!/bin/bash
# ncu captures of the final round-2 code (run under gpurun; outputs in gpurun_out/):
# launch list of the default workload (config 2), --set full of its GEMM and of the few-cells kernel
set -x
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 --clock-sampler off --no-sub-workloads"
$B > gpurun_out/r04c_plain.json 2> gpurun_out/r04c_plain.log || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r04c_launches_cfg2.csv $B > gpurun_out/r04c_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_dense_contract_fp4 -s 2 -c 1 -o gpurun_out/r04c_cfg2_gemm -f $B > gpurun_out/r04c_ncu_cfg2.log 2>&1
ncu --set full --clock-control none -k regex:k_gather_contract -s 2 -c 1 -o gpurun_out/r04c_cfg4_gather -f $B --workload cfg4_bulk > gpurun_out/r04c_ncu_cfg4.log 2>&1
ls -la gpurun_out/
