#!/bin/bash
# same-box A/B of library variants: scripts/gpu_ab_lib.sh <workload> <algo> lib1.so lib2.so ...
mkdir -p gpurun_out
W=$1; A=$2; shift 2
for rep in 1 2; do
for L in "$@"; do
GNNFD_B200_LIB=$PWD/gnn_fraud_detection_b200/$L timeout 300 python bench.py --workload $W --algo $A --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/ab_$L.json 2> gpurun_out/ab_$L.err || { echo "$L failed"; tail -3 gpurun_out/ab_$L.err; }
python -c "
import json; d=json.load(open('gpurun_out/ab_$L.json')); r=d['roofline']; print('$L', round(d['ms_per_step'],2), {k: round(v,2) for k,v in r['stages_ms'].items()})"
done; done
