#!/bin/bash
mkdir -p gpurun_out
for D in 0 1 2 4 3 7; do
GNNFD_DBG=$D timeout 300 python bench.py --workload powerlaw_20m --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_dbg_$D.log 2>&1; echo "DBG=$D $(grep -o '"project_fwd": [0-9.]*' gpurun_out/bench_dbg_$D.log)"
done
