#!/bin/bash
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_partition_gpu.py -x -q -p no:cacheprovider 2>&1 | tail -3
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_scale_2.log 2>&1
echo "G=2 exit $?"; tail -1 gpurun_out/bench_scale_2.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e'])"
