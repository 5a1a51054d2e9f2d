#!/bin/bash
mkdir -p gpurun_out
for L in 0 32 48 64 96 128 0; do
GNNFD_SRC_LOOKAHEAD=$L timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_s_$L.log 2>&1; echo -n "LOOK=$L "; tail -1 gpurun_out/bench_s_$L.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['stages_ms'])"
done
