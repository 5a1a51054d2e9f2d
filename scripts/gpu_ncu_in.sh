#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --workload powerlaw_20m --algo 3 --steps 1 --warmup 3 --no-e2e --no-cpu"
timeout 200 $CMD > gpurun_out/ncu_in_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/ncu_in_plain.log; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gat_in_(fwd|bwd)_items' -s 9 -c 2 -o gpurun_out/prof_in_${1:-a} -f $CMD > gpurun_out/ncu_in_full.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_in_full.log
