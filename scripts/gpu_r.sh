#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider -k "not headline" 2>&1 | tail -2
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_r.log 2>&1; tail -1 gpurun_out/bench_r.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['layer']['frac'], d['roofline']['stages_ms'])"
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/ncu_plain200.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r_200m.csv $CMD > gpurun_out/ncu_launch200.log 2>&1
echo "ncu exit $?"
