#!/bin/bash
# usage: gpu_ng.sh G   -- one scaling point of the default bench (replicate variant), without the e2e leg
mkdir -p gpurun_out
G=$1
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 2951$G bench.py --gpus $G --steps 5 --warmup 3 --no-e2e > gpurun_out/bench_scale_${G}.log 2>&1
echo "G=$G exit $?"; tail -1 gpurun_out/bench_scale_${G}.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['roofline'].get('stages_ms'))"
