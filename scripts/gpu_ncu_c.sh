#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --workload powerlaw_20m --steps 2 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/ncu_c_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'gemm_tc_ws2|dw_tc2|gat_fwd_items|gat_bwd_dst_items|gat_bwd_src_rows|dax_partial' -s 24 -c 6 -o gpurun_out/prof_r1c -f $CMD > gpurun_out/ncu_c_full.log 2>&1
echo "ncu exit $?" >> gpurun_out/ncu_c_full.log
tail -3 gpurun_out/ncu_c_full.log
