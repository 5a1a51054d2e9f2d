#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --workload powerlaw_20m --steps 2 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'gat_(fwd_rows|bwd_dst_rows|bwd_src_rows)' -s 9 -c 3 -o gpurun_out/prof_r1a -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu exit $?" >> gpurun_out/ncu_full.log
$CMD > gpurun_out/ncu_plain2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r1a.csv $CMD > gpurun_out/ncu_launch.log 2>&1
tail -3 gpurun_out/ncu_full.log
