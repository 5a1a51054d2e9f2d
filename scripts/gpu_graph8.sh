#!/bin/bash
mkdir -p gpurun_out
G=${1:-8}
for mode in gather redundant; do
if [ $mode = redundant ]; then export GNNFD_LOGITS_REDUNDANT=1; fi
GNNFD_BENCH_DEBUG=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $G --steps 10 --warmup 3 --no-e2e > gpurun_out/bench_graph_${G}_$mode.log 2> gpurun_out/bench_graph_${G}_$mode.err
echo "G=$G $mode exit $?"; grep "stages_ms" gpurun_out/bench_graph_${G}_$mode.err | cut -c1-330 | head -8; tail -1 gpurun_out/bench_graph_${G}_$mode.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['timing'])"
done
