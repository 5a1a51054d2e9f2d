#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --workload powerlaw_20m --steps 1 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'gemm_tc_ws|dw_tc' -s 6 -c 2 -o gpurun_out/prof_r1d_gemm -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu exit $?"
