#!/bin/bash
mkdir -p gpurun_out
for lib in default alt; do
  if [ $lib = alt ]; then export GNNFD_B200_LIB=$PWD/gnn_fraud_detection_b200/libgnnfd_b200_alt.so; fi
  timeout 300 python bench.py --workload powerlaw_20m --algo 3 --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/ab_$lib.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/ab_$lib.json')); r=d['roofline']; print('$lib', round(d['ms_per_step'],2), {k: round(v,2) for k,v in r['stages_ms'].items()})"
done
