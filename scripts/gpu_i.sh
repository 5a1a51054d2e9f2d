#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "^(FAILED|ERROR)|passed|failed|^E  +(Assert|assert)" gpurun_out/pytest_gpu.log | cut -c1-200 | head -20
timeout 300 python bench.py --workload powerlaw_20m --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_i20m.log 2>&1; grep -o '"ms_per_step": [0-9.]*\|"stages_ms": {[^}]*}' gpurun_out/bench_i20m.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_i200m.log 2>&1; grep -o '"ms_per_step": [0-9.]*\|"stages_ms": {[^}]*}\|"layer": {[^}]*}' gpurun_out/bench_i200m.log
timeout 300 python bench.py --workload elliptic --steps 20 --warmup 5 --no-cpu --no-e2e > gpurun_out/bench_i_ell.log 2>&1; grep -o '"ms_per_step": [0-9.]*\|"stages_ms": {[^}]*}' gpurun_out/bench_i_ell.log
CMD="python bench.py --workload powerlaw_20m --steps 2 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1d.csv $CMD > gpurun_out/ncu_launch.log 2>&1
