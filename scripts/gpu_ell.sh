#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --workload elliptic --steps 2 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/ncu_plain_ell.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_ell.csv $CMD > gpurun_out/ncu_launch_ell.log 2>&1
echo "ncu exit $?"
