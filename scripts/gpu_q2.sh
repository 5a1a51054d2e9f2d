#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider -k "tc or TC or project or gemm" 2>&1 | tail -3
for i in 1 2; do
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_q.log 2>&1; tail -1 gpurun_out/bench_q.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['layer']['frac'], d['roofline']['stages_ms'])"
done
