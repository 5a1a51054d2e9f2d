#!/bin/bash
mkdir -p gpurun_out
for V in old new old new; do
if [ $V = old ]; then export GNNFD_B200_LIB=$PWD/gnn_fraud_detection_b200/libgnnfd_b200_old.so; else unset GNNFD_B200_LIB; fi
GNNFD_SRC_LOOKAHEAD=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_t_$V.log 2>&1; echo -n "$V "; tail -1 gpurun_out/bench_t_$V.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['stages_ms'])"
done
