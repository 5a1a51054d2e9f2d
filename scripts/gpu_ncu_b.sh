#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --workload powerlaw_20m --steps 2 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/ncu_b_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'gemm_tc_ws|dw_tc|gat_bwd_src_rows|dax_partial' -s 12 -c 4 -o gpurun_out/prof_r1b -f $CMD > gpurun_out/ncu_b_full.log 2>&1
echo "ncu exit $?" >> gpurun_out/ncu_b_full.log
tail -3 gpurun_out/ncu_b_full.log
