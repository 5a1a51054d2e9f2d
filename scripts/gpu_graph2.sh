#!/bin/bash
mkdir -p gpurun_out
G=${1:-2}
GNNFD_BENCH_DEBUG=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $G --steps 10 --warmup 3 --no-e2e > gpurun_out/bench_graph_${G}.log 2> gpurun_out/bench_graph_${G}.err
echo "G=$G exit $?"; grep "stages_ms" gpurun_out/bench_graph_${G}.err | cut -c1-420 | head -8; tail -1 gpurun_out/bench_graph_${G}.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['timing'])"
