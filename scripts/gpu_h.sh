#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/ncu_plain200.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'gat_fwd_items' -s 3 -c 1 -o gpurun_out/prof_r1c_200m -f $CMD > gpurun_out/ncu_full200.log 2>&1
echo "ncu exit $?"
$CMD > gpurun_out/ncu_plain200b.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01_200m.csv $CMD > gpurun_out/ncu_launch200.log 2>&1
echo "ncu2 exit $?"
