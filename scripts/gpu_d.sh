#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_d200m.log 2>&1; tail -c 1000 gpurun_out/bench_d200m.log
CMD="python bench.py --workload powerlaw_20m --steps 2 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1b.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "ncu exit $?"
timeout 300 python bench.py --workload elliptic --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_d_elliptic.log 2>&1; tail -c 900 gpurun_out/bench_d_elliptic.log
